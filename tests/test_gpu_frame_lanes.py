"""GPU parity of the frame-lane kernels (alac/net_b200/csrc/kf_frame.cu: one lane per frame and channel from
bitstream to PCM; the default for batches of >= 65536 frames per device) against the oracle.  The scenarios are
the ones of tests/test_gpu_parity.py, re-run with ALACGPU_FLAG_FORCE_FRAME_LANES so that small inputs take the
big-batch path, plus a batch that is big enough to take it by default."""
import numpy as np
import pytest

import tests.test_gpu_parity as P

pytestmark = pytest.mark.gpu

FORCE = 0x80


@pytest.fixture(autouse=True)
def _force_frame_lanes(monkeypatch):
    monkeypatch.setattr(P, "_EXTRA_FLAGS", FORCE)


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("k,scale", [(1, 0.1), (2, 0.02), (3, 0.2)])
def test_configs_small(k, scale, resident, gen, oracle):
    P.test_configs_small(k, scale, resident, gen, oracle)


def test_config1_and_config3_full_size(gen, oracle):
    P.test_config1_full_size(gen, oracle)
    P.test_config3_full_size_divergence(gen, oracle)


def test_config4_and_5_shaped_batches(gen, oracle):
    P.test_config4_and_5_shaped_batches(gen, oracle)


@pytest.mark.parametrize("chunk", [32, 96, 0])
def test_chunking_does_not_change_bytes(chunk, gen, oracle):
    P.test_chunking_and_lane_options_do_not_change_bytes(32, chunk, gen, oracle)


def test_truncated_and_malformed_frames(gen, oracle):
    P.test_truncated_and_malformed_frames_follow_the_oracle_policy(gen, oracle)


def test_cookie_variants(gen, oracle):
    P.test_kmodifier_and_history_cookie_variants(gen, oracle)


def test_extreme_frame_sizes(gen, oracle):
    P.test_extreme_frame_sizes(gen, oracle)


def test_degenerate_signals(gen, oracle):
    P.test_degenerate_signals(gen, oracle)


def test_container_channel_mismatch(gen, oracle):
    P.test_container_channel_mismatch(gen, oracle)


def test_empty_and_ragged_inputs(gen, oracle):
    from alac.net_b200 import BatchDecoder
    t = gen.make_config(1, scale=0.02)[0]
    ref = P._oracle(oracle, t)[0]
    with BatchDecoder(devices=[0], flags=FORCE) as dec:
        dec.add_track(t.cfg, b"", np.zeros(0, np.uint32))
        dec.add_track(t.cfg, t.mdat, t.stsz)
        dec.add_track(t.cfg, b"", np.zeros(0, np.uint32))
        out, off, ln, status = dec.decode_all()
        assert list(ln) == [0, len(ref), 0]
        assert out[int(off[1]):int(off[1] + ln[1])].tobytes() == ref


def test_tracks_added_after_a_decode(gen, oracle):
    P.test_tracks_added_after_a_decode_never_go_back_to_host_memory(gen, oracle)
    P.test_pageable_destination_gets_zero_gaps(gen, oracle)


@pytest.mark.parametrize("seed", range(2))
def test_random_payload_fuzz(seed, gen, oracle):
    from tests.test_fuzz_oracle_model import random_frame
    rng = np.random.default_rng(9000 + seed)
    tracks = []
    for ss, cch, max_n, hm, ih, kmod in ((16, 2, 200, 40, 10, 14), (24, 2, 64, 40, 10, 14), (16, 1, 200, 255, 255, 20),
                                          (24, 1, 16, 4, 0, 6), (16, 2, 4096, 40, 10, 14)):
        frames = [random_frame(rng, ss, max_n)[0] for _ in range(150 if max_n < 4096 else 12)]
        cfg = gen.TrackCfg(ss, cch, max_n, hm, ih, kmod, 44100)
        stsz = np.array([len(f) for f in frames], dtype=np.uint32)
        tracks.append(gen.Track(cfg, b"".join(frames), stsz, np.zeros(len(frames), np.int32), b""))
    for resident in (False, True):
        got, status, _ = P._decode(tracks, flags=0, resident=resident)
        P._assert_tracks_equal(tracks, got, status, oracle, check_encoder=False)
        assert (status == 0).any() and (status != 0).any()


def test_default_policy_takes_the_frame_lanes_for_a_big_batch(gen, monkeypatch):
    """At or above the frame threshold (650 k frames per device; lowered here through ALACGPU_KF_MIN so that the
    test stays small) no flag is needed.  70,000 frames of 256 samples (16-bit stereo and the
    16/24-bit mono/stereo mix), checked against the encoder's input and by the device checksum; the same batch
    through the stream-lane kernels (ALACGPU_FLAG_NO_FRAME_LANES) gives the same bytes."""
    from alac.net_b200 import BatchDecoder, host_checksum
    monkeypatch.setattr(P, "_EXTRA_FLAGS", 0)
    monkeypatch.setenv("ALACGPU_KF_MIN", "65536")
    rng = np.random.default_rng(5)
    tracks = []
    for i, (ss, ch, stereo) in enumerate(((16, 2, True), (16, 1, False), (24, 2, True), (24, 1, False), (16, 2, True))):
        cfg = gen.TrackCfg(ss, ch, 256, 40, 10, 14, 44100)
        n = 256 * 14000 + 100
        x = gen.make_signal(1000 + i, n, ss, 44100, ch, wasted_spans=(ss == 24))
        fr = gen.make_frames(rng, cfg, n, stereo, orders=(0, 31), quants=(1, 15), rice_mods=(1, 7))
        if ss == 24:
            gen.assign_wasted_bytes(fr, x, 24, rng)
        tracks.append(gen.build_track(cfg, x, fr))
    assert sum(t.n_frames for t in tracks) >= 65536
    outs = []
    for flags in (0, 0x40):
        with BatchDecoder(devices=[0], flags=flags) as dec:
            for t in tracks:
                dec.add_track(t.cfg, t.mdat, t.stsz)
            out, off, ln, status = dec.decode_all()
            tm = dec.timing()
            assert (status == 0).all()
            for t, o, l in zip(tracks, off, ln):
                assert out[int(o):int(o + l)].tobytes() == t.pcm
            dec.reindex()
            dec.decode_all(False, want_status=False)
            assert dec.checksum() == host_checksum(out[:dec.total_pcm_bytes()])
            outs.append((out.tobytes(), tm["kernel_launches"], tm["chunks"]))
    assert outs[0][0] == outs[1][0]
    # frame-lane chunks launch K0 + 3 sort kernels + phase A + phase B + pack-only + fix-up = 8 kernels each
    assert outs[0][1] == outs[0][2] * 8
