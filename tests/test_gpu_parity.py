"""GPU parity: libalacgpu (through the C ABI) vs the oracle, byte for byte."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# OR-ed into the context flags of every _decode call: tests/test_gpu_frame_lanes.py re-runs the scenarios of
# this file with ALACGPU_FLAG_FORCE_FRAME_LANES (0x80)
_EXTRA_FLAGS = 0


def _decode(tracks, dst=None, resident=False, **kw):
    """resident=False: decode_all streams the mdat in chunk by chunk (entropy + LPC fused, K3 apart);
    resident=True: prepare() first, and (unless the test chose its own flags) force the fully fused
    launch -- entropy + LPC + pack roles -- which the library otherwise keeps for zero-copy output."""
    from alac.net_b200 import BatchDecoder
    if resident and "flags" not in kw:
        kw["flags"] = 0x20
    kw["flags"] = kw.get("flags", 0) | _EXTRA_FLAGS
    with BatchDecoder(**kw) as dec:
        for t in tracks:
            dec.add_track(t.cfg, t.mdat, t.stsz)
        if resident:
            dec.prepare()
        pcm, off, ln, status = dec.decode_all(dst)
        timing = dec.timing()
        return [pcm[int(o):int(o + l)].tobytes() for o, l in zip(off, ln)], status, timing


def _oracle(oracle, t):
    return oracle.decode_track(oracle.cfg_from(t.cfg), t.mdat, t.stsz)


def _assert_tracks_equal(tracks, got, status, oracle, check_encoder=True):
    pos = 0
    for i, (t, g) in enumerate(zip(tracks, got)):
        ref, st, fb = _oracle(oracle, t)
        if check_encoder:
            assert ref == t.pcm, "oracle does not reproduce the encoder's input"
        assert np.array_equal(status[pos:pos + t.n_frames], st), (status[pos:pos + t.n_frames], st)
        assert len(g) == len(ref)
        if g != ref:
            a, b = np.frombuffer(g, np.uint8), np.frombuffer(ref, np.uint8)
            bad = np.nonzero(a != b)[0]
            raise AssertionError(f"track {i}: {bad.size} bytes differ, first at {bad[:8]}")
        pos += t.n_frames


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("k,scale", [(1, 0.1), (2, 0.02), (3, 0.2)])
def test_configs_small(k, scale, resident, gen, oracle):
    tracks = gen.make_config(k, scale=scale)
    got, status, timing = _decode(tracks, resident=resident)
    _assert_tracks_equal(tracks, got, status, oracle)
    assert timing["kernel_launches"] > 0


def test_config1_full_size(gen, oracle):
    """configs[0] at its full 60 s: 646 frames, last one partial (hassize)"""
    tracks = gen.make_config(1, scale=1.0)
    assert tracks[0].n_frames == 646 and tracks[0].frame_samples[-1] == 4080
    got, status, _ = _decode(tracks)
    _assert_tracks_equal(tracks, got, status, oracle)


def test_config3_full_size_divergence(gen, oracle):
    tracks = gen.make_config(3, scale=1.0)
    got, status, _ = _decode(tracks)
    _assert_tracks_equal(tracks, got, status, oracle)


def test_config4_and_5_shaped_batches(gen, oracle):
    """many tracks in one decode_all: 16-bit stereo batch (config 4) and the 16/24 mono/stereo mix (config 5)"""
    tracks = gen.make_config(4, scale=0.004, n_tracks=24) + gen.make_config(5, scale=0.004, n_tracks=20)
    for resident in (False, True):
        got, status, _ = _decode(tracks, resident=resident)
        _assert_tracks_equal(tracks, got, status, oracle)


@pytest.mark.parametrize("lanes", [8, 16, 32])
@pytest.mark.parametrize("chunk", [32, 96, 0])
def test_chunking_and_lane_options_do_not_change_bytes(lanes, chunk, gen, oracle):
    tracks = gen.make_config(2, scale=0.01) + gen.make_config(3, scale=0.05)
    got, status, _ = _decode(tracks, chunk_frames=chunk, entropy_lanes=lanes)
    _assert_tracks_equal(tracks, got, status, oracle)


@pytest.mark.parametrize("flags", [0, 2, 4])
def test_fused_and_two_kernel_paths_agree(flags, gen, oracle):
    """ALACGPU_FLAG_NO_FUSION (2): entropy and LPC as two launches; default: one fused launch in which
    the LPC warps consume residuals while the entropy lanes are still decoding"""
    tracks = gen.make_config(2, scale=0.02) + gen.make_config(1, scale=0.1) + gen.make_config(3, scale=0.1)
    got, status, tm = _decode(tracks, flags=flags | 0x20, resident=True)
    _assert_tracks_equal(tracks, got, status, oracle)
    # inputs resident: K123 + fix (default) / K12 + K3 (4) / K1 + K2 + K3 (2); K0 + sort ran in prepare / per chunk
    got, status, tm = _decode(tracks, flags=flags)
    _assert_tracks_equal(tracks, got, status, oracle)
    # per chunk (inputs streamed in): K0 + sort + K12 + K3 (default and 4) / K0 + sort + K1 + K2 + K3 (2)
    assert tm["kernel_launches"] == tm["chunks"] * (5 if flags == 2 else 4)


def test_zero_copy_output_into_pinned_memory(gen, oracle):
    """page-locked destination: the pack warps of the fused launch write the PCM straight into it
    (no device PCM, no D2H stage); pageable destination and ALACGPU_FLAG_NO_ZERO_COPY take the copy path"""
    from alac.net_b200 import BatchDecoder, PinnedBuffer
    tracks = gen.make_config(2, scale=0.02) + gen.make_config(3, scale=0.1) + gen.make_config(1, scale=0.05)
    refs = [_oracle(oracle, t)[0] for t in tracks]
    for flags in (0, 8):
        with BatchDecoder(devices=[0], flags=flags) as dec:
            for t in tracks:
                dec.add_track(t.cfg, t.mdat, t.stsz)
            dec.prepare()                            # inputs resident: the fully fused launch + zero-copy output
            buf = PinnedBuffer(dec.total_pcm_bytes())
            buf.array[:] = 0xAB                      # alignment gaps must come back as zeros
            out, off, ln, status = dec.decode_all(buf)
            tm = dec.timing()
            assert (status == 0).all()
            for r, o, l in zip(refs, off, ln):
                assert out[int(o):int(o + l)].tobytes() == r
            for (o, l), nxt in zip(zip(off, ln), list(off[1:]) + [dec.total_pcm_bytes()]):
                assert not out[int(o + l):int(nxt)].any()
            assert (tm["d2h_ms"] == 0.0) == (flags == 0)
            # the per-frame pull still works afterwards (it decodes device-resident on demand)
            assert dec.read_frame(0, 1) == refs[0][24576:2 * 24576]
            buf.free()


def test_resident_path_equals_streaming_path(gen, oracle):
    """prepare() + decode_all (inputs resident) vs decode_all alone (streams H2D chunk by chunk)"""
    from alac.net_b200 import BatchDecoder, host_checksum
    tracks = gen.make_config(2, scale=0.01)
    ref = _oracle(oracle, tracks[0])[0]
    with BatchDecoder(devices=[0]) as dec:
        dec.add_track(tracks[0].cfg, tracks[0].mdat, tracks[0].stsz)
        total = dec.prepare()
        assert total == len(ref)
        a = dec.decode_all()[0].tobytes()
        dec.reindex()
        dec.decode_all(False, want_status=False)          # device-resident
        assert dec.checksum() == host_checksum(ref)
        b = dec.decode_all()[0].tobytes()
    assert a == ref and b == ref


def test_read_frame_is_one_alaccontext_read(gen, oracle):
    from alac.net_b200 import BatchDecoder
    t = gen.make_config(1, scale=0.05)[0]
    ref, st, fbytes = _oracle(oracle, t)
    with BatchDecoder(devices=[0]) as dec:
        tid = dec.add_track(t.cfg, t.mdat, t.stsz)
        assert dec.frame_count(tid) == t.n_frames
        pos = 0
        for f in range(t.n_frames):
            b = dec.read_frame(tid, f)
            assert len(b) == int(fbytes[f]) and b == ref[pos:pos + len(b)]
            assert dec.frame_samples(tid, f) == int(t.frame_samples[f])
            assert dec.frame_status(tid, f) == 0
            pos += len(b)
        assert dec.read_frame(tid, t.n_frames) == b""          # past the end: 0 bytes (AlacContext.cs:182-186)
        assert pos == len(ref)


def test_empty_and_ragged_inputs(gen, oracle):
    from alac.net_b200 import BatchDecoder
    t = gen.make_config(1, scale=0.02)[0]
    ref = _oracle(oracle, t)[0]
    with BatchDecoder(devices=[0]) as dec:
        out, off, ln, status = dec.decode_all()                # no tracks at all
        assert ln.size == 0
        dec.add_track(t.cfg, b"", np.zeros(0, np.uint32))      # a track with no frames
        dec.add_track(t.cfg, t.mdat, t.stsz)
        dec.add_track(t.cfg, b"", np.zeros(0, np.uint32))
        out, off, ln, status = dec.decode_all()
        assert list(ln) == [0, len(ref), 0]
        assert out[int(off[1]):int(off[1] + ln[1])].tobytes() == ref


def test_truncated_and_malformed_frames_follow_the_oracle_policy(gen, oracle):
    """stsz pointing past the mdat (short read), zero-length frames, bad tags, garbage bits"""
    rng = np.random.default_rng(11)
    t = gen.make_config(1, scale=0.03)[0]
    # (a) mdat cut in the middle of a frame
    cut = int(np.cumsum(t.stsz)[t.n_frames // 2] - 100)
    ta = gen.Track(t.cfg, t.mdat[:cut], t.stsz, t.frame_samples, b"")
    # (b) random garbage frames of assorted sizes incl. 0 and 1 byte
    sizes = np.array([0, 1, 2, 7, 64, 300, 2000, 0, 9000], dtype=np.uint32)
    garbage = rng.integers(0, 256, size=int(sizes.sum()), dtype=np.uint8)
    garbage[0 if sizes[0] else 1] &= 0x3F       # keep a few tags valid so the entropy kernel runs on noise
    tb = gen.Track(t.cfg, garbage.tobytes(), sizes, np.zeros(sizes.size, np.int32), b"")
    # (c) valid frames whose element tag is flipped to 2..7
    md = bytearray(t.mdat)
    offs = np.concatenate([[0], np.cumsum(t.stsz.astype(np.int64))])
    for f in range(0, t.n_frames, 3):
        md[offs[f]] = (md[offs[f]] & 0x1F) | (int(rng.integers(2, 8)) << 5)
    tc = gen.Track(t.cfg, bytes(md), t.stsz, t.frame_samples, b"")
    tracks = [ta, tb, tc]
    for resident in (False, True):
        got, status, _ = _decode(tracks, resident=resident)
        _assert_tracks_equal(tracks, got, status, oracle, check_encoder=False)
        assert (status != 0).any()


def test_kmodifier_and_history_cookie_variants(gen, oracle):
    """non-default cookie parameters, incl. k above 16 (two-part Readbits), tiny and zero kmodifier, and a
    history multiplier below 4 (riceHistoryMult / 4 == 0: the history only decays)"""
    rng = np.random.default_rng(21)
    tracks = []
    for kmod, hm, ih in ((14, 40, 10), (6, 63, 200), (20, 40, 10), (2, 20, 0), (23, 255, 255), (31, 40, 10), (1, 40, 10),
                         (0, 40, 10), (3, 4, 0), (14, 3, 10),      # k == 0 everywhere; history multiplier 0
                         (0, 255, 255)):                           # the history goes negative: FS_HISTORY frames
        cfg = gen.TrackCfg(16, 2, 1024, hm, ih, kmod, 44100)
        n = 1024 * 6 + 100
        x = gen.make_signal(int(rng.integers(1, 1 << 30)), n, 16, 44100, 2).copy()
        x[:, ::5] = rng.integers(-32768, 32768, size=x[:, ::5].shape)
        fr = gen.make_frames(rng, cfg, n, True, orders=(0, 31), quants=(0, 15), rice_mods=(0, 7), auto_escape=False)
        tracks.append(gen.build_track(cfg, x, fr))
    got, status, _ = _decode(tracks)
    n_ok = sum(t.n_frames for t in tracks[:-1])
    _assert_tracks_equal(tracks[:-1], got[:-1], status[:n_ok], oracle)
    _assert_tracks_equal(tracks[-1:], got[-1:], status[n_ok:], oracle, check_encoder=False)
    assert (status[n_ok:] == 6).any()


def test_extreme_frame_sizes(gen, oracle):
    """maximum samples per frame (16384: AlacFile.cs:28; order 0 above 4096 samples is the reference's
    Array.Copy fault, status 8) and frames of 1, 3 and 5 samples"""
    rng = np.random.default_rng(7)
    tracks = []
    for ss, ch, msf in ((16, 2, 16384), (24, 2, 8192), (16, 1, 16384), (16, 2, 1), (16, 2, 3), (24, 1, 5)):
        cfg = gen.TrackCfg(ss, ch, msf, 40, 10, 14, 44100)
        n = msf * 3 + max(1, msf // 3)
        x = gen.make_signal(int(rng.integers(1, 1 << 30)), n, ss, 44100, ch).copy()
        fr = gen.make_frames(rng, cfg, n, ch == 2, orders=(0, 31), quants=(0, 15), rice_mods=(0, 7))
        tracks.append(gen.build_track(cfg, x, fr))
    for resident in (False, True):
        got, status, _ = _decode(tracks, resident=resident)
        _assert_tracks_equal(tracks, got, status, oracle, check_encoder=False)


def test_degenerate_signals(gen, oracle):
    """digital silence (one zero run spans the frame, run lengths above the 16-bit escape), full-scale noise
    kept compressed (nearly every symbol takes the nine-ones escape), DC, full-scale alternation, sparse
    impulses"""
    rng = np.random.default_rng(11)
    tracks = []
    for name, ss, ch in (("silence", 16, 2), ("silence", 24, 1), ("noise", 16, 2), ("noise", 24, 2), ("dc", 16, 2),
                         ("alt", 24, 2), ("sparse", 16, 1)):
        cfg = gen.TrackCfg(ss, ch, 4096, 40, 10, 14, 44100)
        n = 4096 * 4 + 777
        lim = 1 << (ss - 1)
        if name == "silence":
            x = np.zeros((ch, n), dtype=np.int32)
        elif name == "noise":
            x = rng.integers(-lim, lim, size=(ch, n)).astype(np.int32)
        elif name == "dc":
            x = np.full((ch, n), lim - 1, dtype=np.int32)
        elif name == "alt":
            x = np.tile(np.array([lim - 1, -lim], dtype=np.int32), (ch, (n + 1) // 2))[:, :n].copy()
        else:
            x = np.zeros((ch, n), dtype=np.int32)
            idx = rng.integers(0, n, size=40)
            x[:, idx] = rng.integers(-lim, lim, size=(ch, 40))
        fr = gen.make_frames(rng, cfg, n, ch == 2, orders=(0, 31), quants=(0, 15), rice_mods=(0, 7),
                             auto_escape=name != "noise")
        tracks.append(gen.build_track(cfg, x, fr))
    for resident in (False, True):
        got, status, _ = _decode(tracks, resident=resident)
        _assert_tracks_equal(tracks, got, status, oracle)


def test_container_channel_mismatch(gen, oracle):
    """mono elements in a 2-channel container (zero-filled right) and stereo elements in a
    1-channel container (left only) -- AlacFile.cs:534-540, :353-354"""
    rng = np.random.default_rng(31)
    tracks = []
    for ss in (16, 24):
        for cch, stereo in ((2, False), (1, True)):
            cfg = gen.TrackCfg(ss, cch, 4096, 40, 10, 14, 44100)
            n = 4096 * 3 + 77
            x = gen.make_signal(int(rng.integers(1, 1 << 30)), n, ss, 44100, 2 if stereo else 1, wasted_spans=(ss == 24))
            fr = gen.make_frames(rng, cfg, n, stereo, escape_prob=0.25)
            if ss == 24:
                gen.assign_wasted_bytes(fr, x, 24, rng)
            tracks.append(gen.build_track(cfg, x, fr))
    got, status, _ = _decode(tracks)
    _assert_tracks_equal(tracks, got, status, oracle)


def test_unsupported_parameters_fail_loudly(gen):
    from alac.net_b200 import BatchDecoder, AlacGpuError
    with BatchDecoder(devices=[0]) as dec:
        for ss, ch in ((20, 2), (32, 2), (8, 1), (16, 3)):
            with pytest.raises(AlacGpuError) as e:
                dec.add_track(gen.TrackCfg(ss, ch, 4096, 40, 10, 14, 44100), b"\0" * 8, np.array([8], np.uint32))
            assert e.value.code == -5
        t = gen.make_config(1, scale=0.01)[0]
        dec.add_track(t.cfg, t.mdat, t.stsz)
        small = np.zeros(16, dtype=np.uint8)
        with pytest.raises(AlacGpuError) as e:
            dec.decode_all(small)
        assert e.value.code == -6


def test_full_size_roundtrip_properties_config2(gen):
    """BASELINE config 2 at full size (14,063 frames, 345.6 MB PCM): the oracle would take a while,
    so check size-independent properties: decode(encode(pcm)) == pcm, and the device checksum of the
    resident PCM equals the checksum of the bytes copied to the host (a checksum of checksums over
    two disjoint halves equals the whole)."""
    from alac.net_b200 import BatchDecoder, host_checksum
    t = gen.make_config(2, scale=1.0)[0]
    assert t.n_frames == 14063 and len(t.pcm) == 345_600_000
    with BatchDecoder(devices=[0]) as dec:
        dec.add_track(t.cfg, t.mdat, t.stsz)
        out, off, ln, status = dec.decode_all()
        assert (status == 0).all()
        assert out[:int(ln[0])].tobytes() == t.pcm
        half = (len(t.pcm) // 2) & ~7
        whole = dec.checksum()
        assert whole == host_checksum(out[:int(ln[0])])
        assert (dec.checksum(0, half) + dec.checksum(half, len(t.pcm) - half)) % (1 << 64) == whole


def test_virtual_device_partition_stitches(gen, oracle):
    """the multi-GPU plan on one GPU: decode each shard of a 4-way partition separately and stitch"""
    from alac.net_b200 import BatchDecoder
    from alac.net_b200.shard import rank_slices
    tracks = gen.make_config(1, scale=0.1) + gen.make_config(3, scale=0.1)
    refs = [_oracle(oracle, t)[0] for t in tracks]
    pieces = {i: [] for i in range(len(tracks))}
    for r in range(4):
        sl = rank_slices([t.stsz for t in tracks], 4, r)
        if not sl:
            continue
        with BatchDecoder(devices=[0]) as dec:
            for s in sl:
                t = tracks[s.track]
                dec.add_track(t.cfg, t.mdat[s.byte_lo:s.byte_hi], t.stsz[s.frame_lo:s.frame_hi])
            out, off, ln, status = dec.decode_all()
            assert (status == 0).all()
            for s, o, l in zip(sl, off, ln):
                pieces[s.track].append((s.frame_lo, out[int(o):int(o + l)].tobytes()))
    for i, ref in enumerate(refs):
        assert b"".join(p for _, p in sorted(pieces[i])) == ref


def test_two_devices_in_one_context(gen, oracle):
    """frame-range partition over 2 GPUs inside one alacgpu context (skipped on a 1-GPU box)"""
    import ctypes as C
    from alac.net_b200 import BatchDecoder, _native as N, host_checksum
    n = C.c_int32(0)
    N.load().alacgpu_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    tracks = gen.make_config(1, scale=0.1) + gen.make_config(2, scale=0.01) + gen.make_config(3, scale=0.1)
    with BatchDecoder(devices=[0, 1]) as dec:
        for t in tracks:
            dec.add_track(t.cfg, t.mdat, t.stsz)
        pcm, off, ln, status = dec.decode_all()
        got = [pcm[int(o):int(o + l)].tobytes() for o, l in zip(off, ln)]
        _assert_tracks_equal(tracks, got, status, oracle)
        # device-resident shards: checksum over both devices == checksum of the host copy
        dec.decode_all(False, want_status=False)
        assert dec.checksum() == host_checksum(pcm[:dec.total_pcm_bytes()])
        f = tracks[1].n_frames - 1
        assert dec.read_frame(1, f) == got[1][len(got[1]) - tracks[1].frame_samples[f] * 6:]


@pytest.mark.parametrize("env", [
    {"ALACGPU_QUAD_MIN_LAST": "3", "ALACGPU_QUAD_MIN_FIRST": "3", "ALACGPU_LPC_WIDE": "0"},   # four-lane LPC for (almost) every stream: T = 2..8
    {"ALACGPU_QUAD_MIN_LAST": "2", "ALACGPU_QUAD_MIN_FIRST": "2", "ALACGPU_LPC_WIDE": "1"},   # eight lanes per stream: T = 1..4
    {"ALACGPU_SMALL_BATCH_FRAMES": "0"},                               # the mid-size thresholds on small inputs
    {"ALACGPU_QUAD_MIN_LAST": "9", "ALACGPU_QUAD_MIN_FIRST": "13"},
    {"ALACGPU_QUAD_MIN_LAST": "0", "ALACGPU_QUAD_MIN_FIRST": "0", "ALACGPU_NO_TAPER": "1"},   # one lane per stream only
    {"ALACGPU_TEST_INJECT_INTERNAL": "1"},      # pretend a fused hand-off timed out: the batch is decoded again unfused
])
def test_tuning_environment_does_not_change_bytes(env):
    """The multi-lane LPC thresholds and the chunk taper are per-process environment knobs."""
    import os, subprocess, sys
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "_env_case.py")], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "env case ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_two_devices_with_the_big_chunk_launch_shape():
    """the two-blocks-per-SM launch (dynamic shared memory above 48 KB) needs its function attribute on
    EVERY device of the context"""
    import ctypes as C
    import os, subprocess, sys
    from alac.net_b200 import _native as N
    n = C.c_int32(0)
    N.load().alacgpu_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    e = dict(os.environ, ALACGPU_SMEM_PAD="44")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "_multidev_case.py"), "0", "1"], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multidev case ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("seed", range(3))
def test_random_payload_fuzz_matches_oracle(seed, gen, oracle):
    """valid headers + random payload bits (tests/test_fuzz_oracle_model.py): PCM and status words of
    every frame must equal the oracle's, OK or not"""
    from tests.test_fuzz_oracle_model import random_frame
    rng = np.random.default_rng(7000 + seed)
    tracks = []
    for ss, cch, max_n, hm, ih, kmod in ((16, 2, 200, 40, 10, 14), (24, 2, 64, 40, 10, 14), (16, 1, 200, 255, 255, 20),
                                          (24, 1, 16, 4, 0, 6), (16, 2, 4096, 40, 10, 14)):
        frames = [random_frame(rng, ss, max_n)[0] for _ in range(150 if max_n < 4096 else 12)]
        cfg = gen.TrackCfg(ss, cch, max_n, hm, ih, kmod, 44100)
        stsz = np.array([len(f) for f in frames], dtype=np.uint32)
        tracks.append(gen.Track(cfg, b"".join(frames), stsz, np.zeros(len(frames), np.int32), b""))
    for flags, resident in ((0x20, True), (0, False), (2, False), (4, True), (0x10, False)):
        got, status, _ = _decode(tracks, flags=flags, resident=resident)
        _assert_tracks_equal(tracks, got, status, oracle, check_encoder=False)
        assert (status == 0).any() and (status != 0).any()


def test_tracks_added_after_a_decode_never_go_back_to_host_memory(gen, oracle):
    """alacgpu.h: `mdat` is borrowed until the first prepare / decode_all after the add returns.  A track added
    AFTER that is staged on its own; the earlier tracks stay resident in HBM even though the caller has reused
    their buffers (the re-plan used to re-copy every earlier track from its stale host pointer)."""
    from alac.net_b200 import BatchDecoder
    ta = gen.make_config(1, scale=0.05)[0]
    tb = gen.make_config(2, scale=0.01)[0]
    tc = gen.make_config(3, scale=0.05)[0]
    refs = [_oracle(oracle, t)[0] for t in (ta, tb, tc)]
    with BatchDecoder(devices=[0], flags=_EXTRA_FLAGS) as dec:
        a = np.frombuffer(ta.mdat, np.uint8).copy()
        dec.add_track(ta.cfg, a, ta.stsz)
        out, off, ln, _ = dec.decode_all()
        assert out[int(off[0]):int(off[0] + ln[0])].tobytes() == refs[0]
        a[:] = 0xA5                                       # the caller's buffer is its own again
        b = np.frombuffer(tb.mdat, np.uint8).copy()
        dec.add_track(tb.cfg, b, tb.stsz)
        dec.prepare()
        b[:] = 0x5A
        c = np.frombuffer(tc.mdat, np.uint8).copy()
        dec.add_track(tc.cfg, c, tc.stsz)
        out, off, ln, status = dec.decode_all()
        assert (status == 0).all()
        for r, o, l in zip(refs, off, ln):
            assert out[int(o):int(o + l)].tobytes() == r
        assert dec.read_frame(0, 2) == refs[0][2 * 16384:3 * 16384]


def test_pageable_destination_gets_zero_gaps(gen, oracle):
    """the 256-byte alignment gaps between tracks are zero bytes in a pageable destination too (copy path)"""
    from alac.net_b200 import BatchDecoder
    tracks = gen.make_config(1, scale=0.013) + gen.make_config(3, scale=0.017) + gen.make_config(2, scale=0.003)
    with BatchDecoder(devices=[0], flags=_EXTRA_FLAGS, chunk_frames=32) as dec:
        for t in tracks:
            dec.add_track(t.cfg, t.mdat, t.stsz)
        total = dec.total_pcm_bytes()
        dst = np.full(total, 0xAB, dtype=np.uint8)
        out, off, ln, status = dec.decode_all(dst)
        ends = [int(o + l) for o, l in zip(off, ln)]
        starts = [int(o) for o in off[1:]] + [total]
        assert any(s > e for e, s in zip(ends, starts))
        for e, s in zip(ends, starts):
            assert not out[e:s].any()
