"""GPU: the host mirror of the reference API (C++ AlacContext / ALACFileReader over the C ABI)
driven the way the reference's callers drive it (README.md:10, ALACFileReader.cs:89-116,
Program.cs:39-52), checked against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref_frames(oracle, t):
    pcm, st, fbytes = oracle.decode_track(oracle.cfg_from(t.cfg), t.mdat, t.stsz)
    offs = np.concatenate([[0], np.cumsum(fbytes.astype(np.int64))])
    return pcm, [pcm[offs[i]:offs[i + 1]] for i in range(t.n_frames)]


@pytest.mark.parametrize("k,scale", [(1, 0.05), (2, 0.004), (3, 0.05)])
def test_alaccontext_read_loop(k, scale, gen, oracle):
    from alac.net_b200.hostmirror import AlacContext
    t = gen.make_config(k, scale=scale)[0]
    ref, frames = _ref_frames(oracle, t)
    ctx = AlacContext(gen.mux_m4a(t))
    assert ctx.GetSampleRate() == t.cfg.sample_rate
    assert ctx.GetNumChannels() == t.cfg.num_channels
    assert ctx.GetBitsPerSample() == t.cfg.sample_size
    assert ctx.GetBytesPerSample() == t.cfg.sample_size // 8
    assert ctx.GetNumSamples() == t.n_sample_frames
    buf = np.zeros(65546 * 3 * 2, dtype=np.uint8)
    total = 0
    for f in range(t.n_frames):
        n = ctx.Read(buf)
        assert n == len(frames[f]) and buf[:n].tobytes() == frames[f]
        total += int(t.frame_samples[f])
        assert ctx.LastSampleNumber == total
    assert ctx.Read(buf) == 0 and ctx.Read(buf) == 0
    ctx.Dispose()


def test_alaccontext_rejects_bad_headers(gen):
    from alac.net_b200.hostmirror import AlacContext, IOException
    t = gen.make_config(1, scale=0.02)[0]
    with pytest.raises(IOException, match="QuickTime movie headers"):
        AlacContext(gen.mux_m4a(t, mdat_first=True))
    with pytest.raises(IOException):
        AlacContext(b"\x00\x00\x00\x08wide" + gen.mux_m4a(t))
    with pytest.raises(IOException):
        AlacContext(b"")


def test_filereader_rechunks_like_wavestream_read(gen, oracle):
    """arbitrary (offset, count) requests; the concatenation is the whole PCM stream"""
    from alac.net_b200.hostmirror import ALACFileReader
    t = gen.make_config(1, scale=0.05)[0]
    ref, _ = _ref_frames(oracle, t)
    rd = ALACFileReader(gen.mux_m4a(t))
    fmt = rd.WaveFormat
    assert fmt == {"SampleRate": 44100, "BitsPerSample": 16, "Channels": 2, "BlockAlign": 4}
    assert rd.Length == t.n_sample_frames * 4 == len(ref)
    rng = np.random.default_rng(5)
    buf = np.zeros(70000, dtype=np.uint8)
    got = bytearray()
    while True:
        count = int(rng.integers(1, 40000))
        off = int(rng.integers(0, 1000))
        n = rd.Read(buf, off, count)
        got += buf[off:off + n].tobytes()
        if n < count:
            break
    assert bytes(got) == ref
    assert rd.Position == rd.Length
    assert rd.Read(buf, 0, 100) == 0


def _expected_after_seek(t, frames, position, bytes_per_sample, nch):
    """the reference's SetPosition + first Read (AlacContext.cs:262-295, :196-203) in numpy terms"""
    cum = np.concatenate([[0], np.cumsum(t.frame_samples.astype(np.int64))])
    f = int(np.searchsorted(cum, position, side="right") - 1)
    off_ints = int(position - cum[f]) * nch
    fr = frames[f]
    out_bytes = len(fr) - off_ints * bytes_per_sample
    skip = off_ints * (2 if bytes_per_sample == 2 else 1)      # ints are samples (16-bit) or bytes (24-bit)
    return f, fr[skip:skip + max(out_bytes, 0)], int(cum[f + 1])


@pytest.mark.parametrize("k,scale", [(1, 0.05), (2, 0.004)])
def test_setposition_matches_reference_semantics(k, scale, gen, oracle):
    from alac.net_b200.hostmirror import AlacContext
    t = gen.make_config(k, scale=scale)[0]
    ref, frames = _ref_frames(oracle, t)
    bps, nch = t.cfg.sample_size // 8, t.cfg.num_channels
    ctx = AlacContext(gen.mux_m4a(t))
    buf = np.zeros(65546 * 3 * 2, dtype=np.uint8)
    rng = np.random.default_rng(8)
    for position in [0, 1, 4095, 4096, 4097, t.n_sample_frames - 1] + [int(x) for x in rng.integers(0, t.n_sample_frames, 6)]:
        ctx.SetPosition(position)
        f, exp, last = _expected_after_seek(t, frames, position, bps, nch)
        assert ctx.LastSampleNumber == last
        n = ctx.Read(buf)
        assert n == len(exp) and buf[:n].tobytes() == exp, (position, f)
        if f + 1 < t.n_frames:                       # the next Read is the next whole frame
            n = ctx.Read(buf)
            assert buf[:n].tobytes() == frames[f + 1]
    # a position past the end leaves the context where it was (the loops fall through, AlacContext.cs:294)
    ctx.SetPosition(0)
    ctx.SetPosition(t.n_sample_frames + 5)
    n = ctx.Read(buf)
    assert buf[:n].tobytes() == frames[0]


def test_filereader_position_roundtrip_like_the_demo(gen, oracle):
    """AlacNetDemo seeks to the middle (Program.cs:49-51): Position = Length / 2"""
    from alac.net_b200.hostmirror import ALACFileReader
    t = gen.make_config(1, scale=0.05)[0]
    ref, frames = _ref_frames(oracle, t)
    rd = ALACFileReader(gen.mux_m4a(t))
    buf = np.zeros(20000, dtype=np.uint8)
    rd.Read(buf, 0, 10000)
    rd.Position = rd.Length // 2
    pos_samples = (rd.Length // 2) // 4
    n = rd.Read(buf, 0, 16)
    assert buf[:n].tobytes() == ref[pos_samples * 4:pos_samples * 4 + 16]


@pytest.mark.parametrize("kw", [dict(), dict(co64=True, mdat_first=True, large_mdat=True), dict(split_stts=True, chunk_frames=3, gap=0)])
def test_tolerant_demux_decodes_chunked_files(kw, gen, oracle):
    """IsoDemux + alacgpu_add_track_offsets: frames addressed through stsc x stco/co64 x stsz with junk
    between chunks (the reference would read the junk as frame data, AlacContext.cs:194-195)"""
    from alac.net_b200 import BatchDecoder, hostmirror as H
    tracks = [gen.make_config(1, scale=0.05)[0], gen.make_config(2, scale=0.004)[0]]
    with BatchDecoder(devices=[0]) as dec:
        for t in tracks:
            m4a = gen.mux_m4a_ex(t, **kw)
            d = H.iso_demux(m4a)
            dec.add_track_offsets(d["cfg"], m4a, d["offsets"], d["stsz"])
        pcm, off, ln, status = dec.decode_all()
        assert (status == 0).all()
        for t, o, l in zip(tracks, off, ln):
            ref = oracle.decode_track(oracle.cfg_from(t.cfg), t.mdat, t.stsz)[0]
            assert pcm[int(o):int(o + l)].tobytes() == ref


def test_cli_batch_to_wav(tmp_path, gen, oracle):
    """alacgpu_decode: two files in one batch -> two .wav files whose data chunks equal the oracle's PCM"""
    import os, subprocess
    from alac.net_b200 import build
    tracks = [gen.make_config(1, scale=0.03)[0], gen.make_config(2, scale=0.003)[0]]
    paths = []
    for i, t in enumerate(tracks):
        p = tmp_path / f"in{i}.m4a"
        p.write_bytes(gen.mux_m4a_ex(t, co64=bool(i)))
        paths.append(str(p))
    outdir = tmp_path / "out"
    outdir.mkdir()
    r = subprocess.run([build.BIN_CLI, "-o", str(outdir), *paths], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    for i, t in enumerate(tracks):
        wav = (outdir / f"in{i}.wav").read_bytes()
        ref = oracle.decode_track(oracle.cfg_from(t.cfg), t.mdat, t.stsz)[0]
        assert wav[:4] == b"RIFF" and wav[44:] == ref
    # --strict = the reference's grammar: rejects the chunked file, accepts the plain one
    plain = tmp_path / "plain.m4a"
    plain.write_bytes(gen.mux_m4a(tracks[0]))
    r = subprocess.run([build.BIN_CLI, "--strict", "-o", str(tmp_path / "p.wav"), str(plain)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "p.wav").read_bytes()[44:] == oracle.decode_track(oracle.cfg_from(tracks[0].cfg), tracks[0].mdat, tracks[0].stsz)[0]
    r = subprocess.run([build.BIN_CLI, "--strict", "-o", str(tmp_path / "q.wav"), paths[0]], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "QuickTime movie headers" in r.stderr
