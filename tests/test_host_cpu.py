"""CPU: the C-ABI library loads and exports what include/alacgpu.h declares; pure host
logic (partition plan, rank sharding under gloo, container grammar of the C++ mirror,
the synthetic muxer).  No compute call is made here -- there is no GPU and no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from alac.net_b200 import build
    build.build_all()
    return build


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "alacgpu.h")).read()
    return sorted(set(re.findall(r"ALACGPU_API\s+[\w\s\*]+?\b(alacgpu_\w+)\s*\(", hdr)))


def test_header_symbols_are_exported(built):
    from alac.net_b200 import _native as N
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(N.EXPORTS) == declared, "ctypes binding and header disagree"
    out = subprocess.check_output(["nm", "-D", "--defined-only", N.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"libalacgpu.so does not export {missing}"
    stray = [s for s in exported if not s.startswith("alacgpu_")]
    assert not stray, f"non-ABI symbols leak out of libalacgpu.so: {stray[:5]}"


def test_csharp_binding_declares_every_symbol_and_flag():
    """csharp/ cannot be compiled here (no dotnet): keep it in step with the header textually"""
    cs = open(os.path.join(ROOT, "csharp", "AlacNet", "AlacGpuNative.cs")).read()
    hdr = open(os.path.join(ROOT, "include", "alacgpu.h")).read()
    dllimports = sorted(set(re.findall(r"extern\s+\w+\s+(alacgpu_\w+)\s*\(", cs)))
    assert dllimports == _declared_symbols(), "P/Invoke declarations and header disagree"
    flags = {int(v, 16) for v in re.findall(r"#define ALACGPU_FLAG_\w+ (0x[0-9a-fA-F]+)u", hdr)}
    cs_flags = {int(v, 16) for v in re.findall(r"=\s*(0x[0-9a-fA-F]+)", cs[cs.index("enum AlacGpuFlags"):cs.index("struct AlacGpuOpts")])}
    assert flags == cs_flags and len(flags) >= 6
    frame_codes = {int(v) for v in re.findall(r"ALACGPU_FRAME_\w+ = (\d+)", hdr)}
    cs_codes = {int(v) for v in re.findall(r"=\s*(\d+)", cs[cs.index("enum AlacGpuFrameStatus"):cs.index("enum AlacGpuFlags")])}
    assert frame_codes == cs_codes


def test_library_loads_and_answers_without_a_gpu(built):
    from alac.net_b200 import _native as N
    L = N.load()
    assert L.alacgpu_abi_version() == 2
    assert L.alacgpu_strerror(0) == b"ok"
    assert b"no CPU fallback" in L.alacgpu_strerror(N.ERR_NAMES and -2)
    n = C.c_int32(-1)
    assert L.alacgpu_device_count(C.byref(n)) == 0
    if n.value == 0:      # this container: creating a context must fail loudly, never fall back
        h = C.c_void_p()
        assert L.alacgpu_create(None, 0, None, C.byref(h)) == -2
        assert not h.value
        from alac.net_b200 import BatchDecoder, AlacGpuError
        with pytest.raises(AlacGpuError):
            BatchDecoder()


def test_product_never_touches_the_oracle():
    """the product tree must not import / link / mention the oracle or the Python model"""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "alac")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"(from|import)\s+(oracle|pymodel)|alac_oracle|libalac_oracle", src):
                    bad.append(f)
    assert not bad, bad


def test_plan_partition_properties(built):
    from alac.net_b200 import plan_partition
    rng = np.random.default_rng(3)
    sizes = rng.integers(100, 25000, size=10007).astype(np.uint32)
    total = int(sizes.sum())
    for parts in (1, 2, 3, 4, 8):
        cut = plan_partition(sizes, parts)
        assert cut[0] == 0 and cut[-1] == sizes.size and np.all(np.diff(cut.astype(np.int64)) >= 0)
        loads = [int(sizes[int(a):int(b)].sum()) for a, b in zip(cut[:-1], cut[1:])]
        assert sum(loads) == total
        assert max(loads) - min(loads) <= 2 * 25000, loads            # balanced to within a frame or two
    # degenerate inputs
    assert list(plan_partition(np.zeros(0, np.uint32), 4)) == [0, 0, 0, 0, 0]
    assert list(plan_partition(np.array([5], np.uint32), 4))[-1] == 1
    cut = plan_partition(np.array([1, 1, 1000000, 1, 1], np.uint32), 2)
    assert 0 <= cut[1] <= 5


def test_rank_slices_cover_every_frame_once(built):
    from alac.net_b200.shard import rank_slices
    rng = np.random.default_rng(4)
    tracks = [rng.integers(50, 9000, size=n).astype(np.uint32) for n in (17, 1, 300, 0, 45)]
    for world in (1, 2, 4, 8):
        seen = [np.zeros(t.size, dtype=np.int32) for t in tracks]
        for r in range(world):
            for s in rank_slices(tracks, world, r):
                seen[s.track][s.frame_lo:s.frame_hi] += 1
                offs = np.concatenate([[0], np.cumsum(tracks[s.track].astype(np.int64))])
                assert (s.byte_lo, s.byte_hi) == (offs[s.frame_lo], offs[s.frame_hi])
        assert all((x == 1).all() for x in seen)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from alac.net_b200.shard import rank_slices, reduce_step
from oracle import oracle as O            # checker only (this is a test)
from tools.alacgen import alacgen as G
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
tracks = [G.make_config(1, scale=0.02)[0], G.make_config(3, scale=0.03)[0]]
whole = [O.decode_track(O.cfg_from(t.cfg), t.mdat, t.stsz)[0] for t in tracks]
# each rank decodes ONLY its frame ranges (stand-in for its GPU), then the shards are stitched
mine = []
for s in rank_slices([t.stsz for t in tracks], world, rank):
    t = tracks[s.track]
    pcm, st, fb = O.decode_track(O.cfg_from(t.cfg), t.mdat[s.byte_lo:s.byte_hi], t.stsz[s.frame_lo:s.frame_hi])
    mine.append((s.track, s.frame_lo, pcm))
gathered = [None] * world
dist.all_gather_object(gathered, mine)
ms, units = reduce_step(dist, "cpu", 10.0 + rank, float(sum(len(p) for _, _, p in mine)))
if rank == 0:
    parts = sorted((x for g in gathered for x in g), key=lambda x: (x[0], x[1]))
    for ti in range(len(tracks)):
        stitched = b"".join(p for t, _, p in parts if t == ti)
        assert stitched == whole[ti] == tracks[ti].pcm, "stitched shards differ from the whole-track decode"
    assert ms == 10.0 + world - 1 and units == float(sum(len(w) for w in whole))
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_frame_range_sharding_world_size_2_gloo(tmp_path, built, oracle, gen):
    script = tmp_path / "gloo_worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "GLOO_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---- container grammar: the C++ QtMovieT mirror against the synthetic muxer -------------
def test_demux_tables_match_the_muxer(built, gen):
    from alac.net_b200 import hostmirror as H
    for k, scale in ((1, 0.05), (2, 0.005), (3, 0.05)):
        t = gen.make_config(k, scale=scale)[0]
        for kw in ({}, {"free_atom": True}):
            m4a = gen.mux_m4a(t, **kw)
            d = H.demux(m4a)
            assert d["status"] == 1                                        # MdatPosStatus.Ok
            assert (d["sample_size"], d["num_channels"], d["sample_rate"]) == (t.cfg.sample_size, t.cfg.num_channels, t.cfg.sample_rate)
            assert np.array_equal(d["stsz"], t.stsz)
            assert m4a[d["mdat_pos"]:d["mdat_pos"] + len(t.mdat)] == t.mdat
            cd = d["codec_data"]
            assert int(cd[24]) << 24 | int(cd[25]) << 16 | int(cd[26]) << 8 | int(cd[27]) == t.cfg.max_samples_per_frame
            assert (cd[30], cd[31], cd[32]) == (t.cfg.rice_history_mult, t.cfg.rice_initial_history, t.cfg.rice_kmodifier)


def test_demux_rejections_match_the_reference(built, gen):
    from alac.net_b200 import hostmirror as H
    t = gen.make_config(1, scale=0.02)[0]
    good = gen.mux_m4a(t)
    # mdat before moov: SetSavedMdat compares Seek()'s returned position with 0 -> CannotSeek (QTMovieT.cs:744-748)
    assert H.demux(gen.mux_m4a(t, mdat_first=True))["status"] == 3
    # unknown top-level atom -> None (QTMovieT.cs:103-107)
    bad = good[:28] + b"\x00\x00\x00\x08wide" + good[28:]
    assert H.demux(bad)["status"] == 0
    # truncated before mdat -> EOF -> None
    assert H.demux(good[:200])["status"] == 0
    # an extra atom inside stbl (e.g. 'sgpd') is rejected (QTMovieT.cs:221-225)
    i = good.index(b"stco") - 4
    broken = bytearray(good)
    broken[i + 4:i + 8] = b"co64"
    assert H.demux(bytes(broken))["status"] == 0


def test_uniform_stsz(built, gen):
    from alac.net_b200 import hostmirror as H
    cfg = gen.TrackCfg(16, 2, 64, 40, 10, 14, 44100)
    rng = np.random.default_rng(9)
    x = rng.integers(-30000, 30000, size=(2, 64 * 5)).astype(np.int32)
    fr = gen.make_frames(rng, cfg, 64 * 5, True, escape_prob=1.0, end_tag=False)
    t = gen.build_track(cfg, x, fr)
    assert len(set(t.stsz.tolist())) == 1
    d = H.demux(gen.mux_m4a(t, uniform_stsz=True))
    assert d["status"] == 1 and np.array_equal(d["stsz"], t.stsz)


# ---- tolerant demux and the WAV writer ------------------------------------
VARIANTS = [dict(), dict(co64=True), dict(mdat_first=True), dict(split_stts=True, chunk_frames=3, gap=0),
            dict(large_mdat=True, mdat_first=True, co64=True), dict(extra_atoms=False, chunk_frames=1, gap=5)]


@pytest.mark.parametrize("kw", VARIANTS)
def test_iso_demux_resolves_chunked_layouts(kw, built, gen):
    from alac.net_b200 import hostmirror as H
    t = gen.make_config(1, scale=0.03)[0]
    m4a = gen.mux_m4a_ex(t, **kw)
    d = H.iso_demux(m4a)
    cfg = d["cfg"]
    assert (cfg.sample_size, cfg.num_channels, cfg.max_samples_per_frame, cfg.sample_rate) == (16, 2, 4096, 44100)
    assert (cfg.rice_history_mult, cfg.rice_initial_history, cfg.rice_kmodifier) == (40, 10, 14)
    assert np.array_equal(d["stsz"], t.stsz) and np.array_equal(d["durations"], t.frame_samples)
    assert d["total_samples"] == t.n_sample_frames
    offs = np.concatenate([[0], np.cumsum(t.stsz.astype(np.int64))])
    for f in (0, 1, t.n_frames // 2, t.n_frames - 1):
        o = int(d["offsets"][f])
        assert m4a[o:o + int(t.stsz[f])] == t.mdat[offs[f]:offs[f + 1]]
    # the reference's grammar rejects every one of these layouts (unknown atoms, co64, mdat first, ...)
    if kw.get("extra_atoms", True) or kw.get("co64") or kw.get("mdat_first"):
        assert H.demux(m4a)["status"] in (0, 3)
    # and the plain layout parses identically through both demuxers
    plain = gen.mux_m4a(t)
    a, b = H.iso_demux(plain), H.demux(plain)
    assert np.array_equal(a["stsz"], b["stsz"]) and int(a["offsets"][0]) == b["mdat_pos"]


def test_wav_header_fields(built):
    import struct
    from alac.net_b200 import hostmirror as H
    h = H.wav_header(96000, 24, 2, 345_600_000)
    assert h[:4] == b"RIFF" and h[8:16] == b"WAVEfmt " and h[36:40] == b"data"
    riff, = struct.unpack("<I", h[4:8])
    fmt_len, tag, ch, rate, bps, align, bits = struct.unpack("<IHHIIHH", h[16:36])
    data, = struct.unpack("<I", h[40:44])
    assert (fmt_len, tag, ch, rate, bps, align, bits) == (16, 1, 2, 96000, 96000 * 6, 6, 24)
    assert data == 345_600_000 and riff == data + 36


def test_iso_demux_survives_crafted_sizes(built, gen):
    """A 64-bit box size near 2^64 must not wrap the bounds check (next_box), and a uniform stsz must
    not allocate more entries than the file could hold: the demuxer answers (or refuses) without reading
    outside the buffer."""
    import struct
    from alac.net_b200 import hostmirror as H
    t = gen.make_config(1, scale=0.02)[0]
    m4a = bytearray(gen.mux_m4a_ex(t))
    at = bytes(m4a).find(b"sgpd") - 4                     # an ignorable box inside stbl, 8 + 12 bytes
    assert at > 0
    for largesize in (0xFFFFFFFFFFFFFFF0, 0xFFFFFFFFFFFFFFFF - at, 1 << 63, 17):
        bad = bytearray(m4a)
        bad[at:at + 4] = struct.pack(">I", 1)
        bad[at + 8:at + 16] = struct.pack(">Q", largesize)
        try:
            d = H.iso_demux(bytes(bad))
            assert d["stsz"].size <= t.n_frames
        except H.IOException:
            pass
    # uniform stsz: sample_size != 0, absurd count
    plain = bytearray(gen.mux_m4a(t))
    s = bytes(plain).find(b"stsz") + 4
    plain[s + 4:s + 12] = struct.pack(">II", 4096, 0xFFFFFFFF)
    d = H.iso_demux(bytes(plain))
    assert d["stsz"].size <= len(plain) // 4096 + 1


def test_bench_checksum_bookkeeping(gen):
    """bench.py checks a batch of physically replicated tracks against the encoder's input without touching every
    replica: the position-weighted checksum of a track placed at 8-byte word `first` is S1 + 2 first S0 (S0 = sum of
    words, S1 = the checksum at position 0).  Pieces cut by a shard boundary are summed directly."""
    import types
    import bench
    from alac.net_b200 import host_checksum
    args = types.SimpleNamespace(scale=0.004, tracks=7, unique=3, tracks_per_gpu=0, unique_per_kind=1, orders="0,31")
    corpus = bench.make_corpus("config4", 1, args)
    assert corpus.n_tracks == 7 and len(corpus.uniq) == 3
    pieces = bench.rank_pieces(corpus, 1, 0)
    # lay the tracks out like the decoder does (256-byte aligned starts) and compare with the formula
    off, layout = 0, []
    for p in pieces:
        off = (off + 255) // 256 * 256
        layout.append((off, p.p_hi - p.p_lo))
        off += p.p_hi - p.p_lo
    buf = np.zeros(off, dtype=np.uint8)
    for p, (o, l) in zip(pieces, layout):
        buf[o:o + l] = np.frombuffer(corpus.uniq[p.u].pcm, dtype=np.uint8)[p.p_lo:p.p_hi]

    class FakeDecoder:
        def track_pcm_bytes(self, i):
            return layout[i]
    assert bench.expected_checksum(corpus, FakeDecoder(), pieces) == host_checksum(buf)
    # a sharded run: the pieces of both ranks cover every frame once, and partial pieces take the direct sum
    args5 = types.SimpleNamespace(scale=0.004, tracks=20, unique=3, tracks_per_gpu=10, unique_per_kind=1, orders="0,31")
    c5 = bench.make_corpus("config5", 2, args5)
    got = {}
    for rank in range(2):
        for p in bench.rank_pieces(c5, 2, rank):
            got.setdefault(p.j, []).append((p.f_lo, p.f_hi))
    for j in range(c5.n_tracks):
        spans = sorted(got[j])
        assert spans[0][0] == 0 and spans[-1][1] == c5.track(j).n_frames
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    groups = bench.groups_for(bench.rank_pieces(c5, 2, 0), 1 << 20)
    assert sum(len(g) for g in groups) == len(bench.rank_pieces(c5, 2, 0)) and len(groups) > 1
