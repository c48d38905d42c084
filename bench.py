#!/usr/bin/env python
"""bench.py -- decoded PCM Msamples/s of the ALAC frame-decode path on B200.

    python bench.py --gpus 1 --steps K --warmup W             (ours, N=1)
    torchrun ... bench.py --gpus N --steps K --warmup W       (ours, N>1: one rank per GPU)
    python bench.py --impl reference ...                      (the oracle's C port on all host cores)
    python bench.py --gpus N --single-process                 (ONE alacgpu context over N GPUs, one process)

Workloads (BASELINE.json):
  N = 1  configs[3]: 1,000 synthetic 16-bit stereo 44.1 kHz tracks x 252 s (~70 h) decoded in ONE
         alacgpu_decode_all on one B200.  `--unique` distinct tracks are generated and replicated PHYSICALLY
         (every replica is staged separately in HBM: 29.5 GB compressed in, 44.5 GB PCM out per step).
         Secondary `latency_legs`: configs[0], [1], [2] (single tracks, fewer frames than the GPU has lanes).
  N > 1  configs[4]: the 16/24-bit mono/stereo mix (track i is of kind i mod 10), 1,250 tracks per GPU
         (10,000 at N = 8), cut FRAME-WISE into N contiguous ranges balanced by compressed bytes
         (alac.net_b200.shard.rank_slices == alacgpu_plan_partition); every rank stages and decodes only its
         range.  No data-path collective: NCCL carries the barrier and the max/sum of the timings only.

A "step" is one pass of the whole kernel path (header pre-pass + decode kernels) over the workload with the
compressed input already resident in HBM; `value` = channel values decoded by all ranks / max-over-ranks time.
`e2e` is the same metric through the C ABI with HOST buffers: every step clears the context, adds every
track again (H2D of all mdat bytes), decodes and copies all PCM back (D2H), in as few decode_all calls as the
host output buffer allows.  Parity (device checksum of the decoded PCM == checksum of the encoder's input,
frame status all OK, byte compare of the host copy) is checked in the run before anything is timed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decoded PCM Msamples/s"
UNIT = "Msamples/s"
M64 = (1 << 64) - 1


# ---------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------
class Corpus:
    """`uniq`: distinct tracks; `index[j]`: which of them global track j is a (physical) replica of."""

    def __init__(self, uniq, index, desc, key):
        self.uniq, self.index, self.desc, self.key = uniq, np.asarray(index, dtype=np.int64), desc, key
        self._sums = None

    @property
    def n_tracks(self):
        return int(self.index.size)

    def track(self, j):
        return self.uniq[int(self.index[j])]

    def totals(self):
        cnt = np.bincount(self.index, minlength=len(self.uniq))
        f = lambda g: int(sum(int(c) * g(t) for c, t in zip(cnt, self.uniq)))
        return {"frames": f(lambda t: t.n_frames), "samples": f(lambda t: t.n_samples),
                "pcm_bytes": f(lambda t: len(t.pcm)), "compressed_bytes": f(lambda t: len(t.mdat))}

    def sums(self):
        """(S0, S1) of every distinct track's PCM: sum of 8-byte words, and sum of word * (2 j + 1)."""
        if self._sums is None:
            from alac.net_b200 import host_checksum
            out = []
            for t in self.uniq:
                a = np.frombuffer(t.pcm, dtype=np.uint8)
                pad = (-a.size) % 8
                if pad:
                    a = np.concatenate([a, np.zeros(pad, dtype=np.uint8)])
                with np.errstate(over="ignore"):
                    s0 = int(np.sum(a.view("<u8"), dtype=np.uint64))
                out.append((s0, host_checksum(t.pcm)))
            self._sums = out
        return self._sums


def make_corpus(name: str, world: int, args) -> Corpus:
    # torchrun exports OMP_NUM_THREADS=1; the synthetic encoder (OpenMP over frames) should use this rank's share of the host
    ranks_here = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    os.environ["OMP_NUM_THREADS"] = str(max(1, host_cores() // max(1, ranks_here)))
    from tools.alacgen import alacgen as g
    g.build_encoder()
    s = args.scale
    if name == "config1":
        return Corpus([g.track_16_stereo(g.SEED_BASE + 1, 60.0 * s)], [0],
                      "configs[0]: synthetic 16-bit stereo 44.1 kHz ALAC, 60 s, 4096-sample frames", name)
    if name == "config2":
        return Corpus([g.track_24_stereo(g.SEED_BASE + 2, 600.0 * s)], [0],
                      "configs[1]: synthetic 24-bit stereo 96 kHz ALAC, 10 min, 4096-sample frames, LPC order 1..31, "
                      "wasted bytes 0/1/2", name)
    if name == "config3":
        return Corpus([g.track_16_mono_mixed(g.SEED_BASE + 3, 60.0 * s)], [0],
                      "configs[2]: 16-bit mono mixing compressed / uncompressed / Rice-escape frames", name)
    if name == "config4":
        n = args.tracks or 1000
        u = max(1, min(args.unique, n))
        uniq = [g.track_16_stereo(g.SEED_BASE + 1000 + i, 252.0 * s) for i in range(u)]
        return Corpus(uniq, np.arange(n) % u,
                      f"configs[3]: batch of {n} synthetic 16-bit stereo 44.1 kHz tracks x {252.0 * s:g} s decoded in one "
                      f"call ({u} distinct tracks, every replica staged separately in HBM)", name)
    if name == "fixed":
        # tuning runs: 16-bit stereo tracks whose frames all use predictor orders in --orders lo,hi
        lo, hi = (int(x) for x in args.orders.split(","))
        n = args.tracks or 32
        u = max(1, min(args.unique, n))
        uniq = []
        for i in range(u):
            seed = g.SEED_BASE + 5000 + i
            rng = np.random.default_rng(seed)
            cfg = g.TrackCfg(16, 2, 4096, 40, 10, 14, 44100)
            ns = int(round(252.0 * s * 44100))
            x = g.make_signal(seed, ns, 16, 44100, 2)
            uniq.append(g.build_track(cfg, x, g.make_frames(rng, cfg, ns, True, orders=(lo, hi), quants=(1, 15), rice_mods=(1, 7))))
        return Corpus(uniq, np.arange(n) % u, f"tuning: {n} 16-bit stereo tracks x {252.0 * s:g} s, predictor orders {lo}..{hi}", "fixed")
    if name == "config5":
        per = args.tracks_per_gpu or 1250
        n = args.tracks or per * world
        u = 10 * max(1, args.unique_per_kind)
        uniq = [g.corpus_track(i, 252.0 * s) for i in range(u)]
        return Corpus(uniq, np.arange(n) % u,
                      f"configs[4]: {n}-track mixed corpus x {252.0 * s:g} s (track i is of kind i mod 10: 0-4 16-bit stereo "
                      f"44.1 kHz, 5-6 16-bit mono, 7-8 24-bit stereo 48 kHz, 9 24-bit mono 96 kHz; {u} distinct tracks, "
                      f"replicas staged separately), {per} tracks per GPU, cut frame-wise over {world} GPU(s)", name)
    raise SystemExit(f"unknown workload {name}")


class Piece:
    """Frames [f_lo, f_hi) of global track j (a whole track unless a shard boundary cuts it)."""
    __slots__ = ("j", "u", "f_lo", "f_hi", "b_lo", "b_hi", "p_lo", "p_hi")

    def __init__(self, corpus, j, f_lo=None, f_hi=None, b_lo=None, b_hi=None):
        t = corpus.track(j)
        self.j, self.u = j, int(corpus.index[j])
        self.f_lo, self.f_hi = (0, t.n_frames) if f_lo is None else (f_lo, f_hi)
        if b_lo is None:
            offs = np.concatenate([[0], np.cumsum(t.stsz.astype(np.int64))])
            b_lo, b_hi = int(offs[self.f_lo]), int(offs[self.f_hi])
        self.b_lo, self.b_hi = b_lo, b_hi
        bpf = (t.cfg.sample_size // 8) * t.cfg.num_channels
        ps = np.concatenate([[0], np.cumsum(t.frame_samples.astype(np.int64))]) * bpf
        self.p_lo, self.p_hi = int(ps[self.f_lo]), int(ps[self.f_hi])

    def whole(self, corpus):
        return self.f_lo == 0 and self.f_hi == corpus.track(self.j).n_frames


def rank_pieces(corpus: Corpus, world: int, rank: int):
    if world == 1:
        out = []
        cache = {}
        for j in range(corpus.n_tracks):
            u = int(corpus.index[j])
            if u not in cache:
                cache[u] = Piece(corpus, j)
            c = cache[u]
            p = Piece.__new__(Piece)
            p.j, p.u, p.f_lo, p.f_hi, p.b_lo, p.b_hi, p.p_lo, p.p_hi = j, u, c.f_lo, c.f_hi, c.b_lo, c.b_hi, c.p_lo, c.p_hi
            out.append(p)
        return out
    from alac.net_b200.shard import rank_slices
    sl = rank_slices([corpus.track(j).stsz for j in range(corpus.n_tracks)], world, rank)
    return [Piece(corpus, s.track, s.frame_lo, s.frame_hi, s.byte_lo, s.byte_hi) for s in sl]


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.recording = False          # the thread runs through the warm-up (NVML's first queries are slow and were seen
                                        # to stall the launching thread for 0.1-0.5 s); samples count from begin() on
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.recording:
                    self.samples.append(mhz)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def begin(self):
        self.recording = True

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def mem_available() -> int:
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return 64 << 30


# ---------------------------------------------------------------------------
# CPU arm: the oracle's C port on the host cores
# ---------------------------------------------------------------------------
def cpu_decode_rate(corpus: Corpus, frames_per_thread: int, threads: int, repeats: int = 1):
    """One contiguous frame range per thread (frames are independent, so a thread decodes its range exactly
    as one AlacContext would pump it); thread i takes its range from distinct track i mod U, so a mixed corpus
    is sampled in its mix.  -> (Msamples/s, sample text, seconds)."""
    from oracle import oracle as o
    o.build()
    jobs = []
    samples = 0
    for i in range(threads):
        t = corpus.uniq[i % len(corpus.uniq)]
        per = max(1, min(frames_per_thread, t.n_frames))
        a = (i * per) % max(1, t.n_frames - per + 1)
        offs = np.concatenate([[0], np.cumsum(t.stsz.astype(np.int64))])
        mdat = np.frombuffer(t.mdat, dtype=np.uint8)
        jobs.append((o.cfg_from(t.cfg), mdat[offs[a]:offs[a + per]].tobytes(), t.stsz[a:a + per]))
        samples += int(t.frame_samples[a:a + per].sum()) * t.cfg.num_channels

    def work(cfg, data, stsz):
        o.decode_track(cfg, data, stsz)

    best = None
    for _ in range(repeats):
        ths = [threading.Thread(target=work, args=j) for j in jobs]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    text = (f"{threads} threads x {frames_per_thread} frames ({samples} samples) of the workload"
            + (f", thread i on distinct track i mod {len(corpus.uniq)}" if len(corpus.uniq) > 1 else "") + f", best of {repeats}")
    return samples / best / 1e6, text, best


def default_workload(args, world):
    if args.workload != "auto":
        return args.workload
    return "config4" if world == 1 else "config5"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = max(1, args.gpus)
    corpus = make_corpus(default_workload(args, world), world, args)
    cores = host_cores()
    # bounded sample: calibrate on a small slice, then size each step so the whole --steps/--warmup run stays
    # near a 90 s budget (at most 2048 frames per thread and step)
    _, _, dt0 = cpu_decode_rate(corpus, 8, cores)
    per_frame = dt0 / 8.0
    budget = 90.0 / max(1, args.steps + args.warmup)
    per = int(max(8, min(2048, budget / max(per_frame, 1e-6))))
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, sample, dt = cpu_decode_rate(corpus, per, cores)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": corpus.desc, "scale": args.scale},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + " per step; oracle/ C restatement of AlacFile.cs (the C# reference cannot run: no .NET in the image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class Staged:
    """Pinned host copies of the distinct tracks' mdat bytes + helpers to add pieces to a decoder."""

    def __init__(self, corpus: Corpus, pinned: bool = True):
        from alac.net_b200 import PinnedBuffer
        self.corpus = corpus
        self.src, self._keep = [], []
        for t in corpus.uniq:
            if pinned:
                pb = PinnedBuffer(len(t.mdat))
                pb.array[:] = np.frombuffer(t.mdat, dtype=np.uint8)
                self.src.append(pb.array)
                self._keep.append(pb)
            else:
                self.src.append(np.frombuffer(t.mdat, dtype=np.uint8).copy())       # plain pageable memory

    def add(self, dec, p: Piece):
        t = self.corpus.uniq[p.u]
        dec.add_track(t.cfg, self.src[p.u][p.b_lo:p.b_hi], t.stsz[p.f_lo:p.f_hi])


def expected_checksum(corpus: Corpus, dec, pieces) -> int:
    """Checksum of the encoder's input laid out like the decoder's output, from per-track sums:
    sum_j w_j (2 (first + j) + 1) = S1 + 2 first S0 for a whole track placed at 8-byte word `first`."""
    from alac.net_b200 import host_checksum
    sums = corpus.sums()
    total = 0
    for i, p in enumerate(pieces):
        off, ln = dec.track_pcm_bytes(i)
        assert ln == p.p_hi - p.p_lo, (i, ln, p.p_hi - p.p_lo)
        if p.whole(corpus):
            s0, s1 = sums[p.u]
            total += s1 + 2 * (off // 8) * s0
        else:
            total += host_checksum(corpus.uniq[p.u].pcm[p.p_lo:p.p_hi], first_word=off // 8)
    return total & M64


def check_host_copy(corpus, pieces, out, off, ln, every: int = 1) -> bool:
    ok = True
    for i in range(0, len(pieces), every):
        p = pieces[i]
        ref = np.frombuffer(corpus.uniq[p.u].pcm, dtype=np.uint8)[p.p_lo:p.p_hi]
        ok = ok and int(ln[i]) == ref.size and np.array_equal(out[int(off[i]):int(off[i] + ln[i])], ref)
    return ok


def groups_for(pieces, cap_bytes):
    """consecutive pieces whose PCM (with 256-byte alignment per piece) fits cap_bytes"""
    out, cur, used = [], [], 0
    for p in pieces:
        need = (p.p_hi - p.p_lo + 255) // 256 * 256
        if cur and used + need > cap_bytes:
            out.append(cur)
            cur, used = [], 0
        cur.append(p)
        used += need
    if cur:
        out.append(cur)
    return out


def e2e_leg(dec, staged, pieces, steps, warm, barrier, pinned_out: bool, cap_bytes: int):
    """clear -> add every piece (H2D of all mdat bytes) -> decode_all into a HOST buffer (D2H of all PCM), per
    group of pieces that fits the output buffer.  -> (ms per step, groups, last timing, parity ok)."""
    from alac.net_b200 import PinnedBuffer
    groups = groups_for(pieces, cap_bytes)
    need = max(sum((p.p_hi - p.p_lo + 255) // 256 * 256 for p in g) for g in groups)
    if pinned_out:
        buf = PinnedBuffer(need)
        arr = buf.array
    else:
        buf = None
        arr = np.empty(need, dtype=np.uint8)
    ok = True

    def step(check=False):
        nonlocal ok
        for g in groups:
            dec.clear()
            for p in g:
                staged.add(dec, p)
            out, off, ln, _ = dec.decode_all(arr, want_status=False)
            if check:
                ok = ok and check_host_copy(staged.corpus, g, out, off, ln, every=max(1, len(g) // 24))

    step(check=True)
    for _ in range(max(0, warm - 1)):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    barrier()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    tm = dec.timing()
    if buf is not None:
        buf.free()
    return ms, len(groups), tm, ok


def latency_leg(name, args, local):
    """A single-track config (fewer frames than lanes): device-resident and end-to-end time of one decode."""
    from alac.net_b200 import BatchDecoder, PinnedBuffer, host_checksum
    corpus = make_corpus(name, 1, args)
    t = corpus.uniq[0]
    staged = Staged(corpus)
    with BatchDecoder(devices=[local], flags=args.flags) as dec:
        staged.add(dec, Piece(corpus, 0))
        total = dec.prepare()
        host = PinnedBuffer(total)
        out, off, ln, status = dec.decode_all(host)
        ok = bool((status == 0).all()) and out[:int(ln[0])].tobytes() == t.pcm
        dec.decode_all(False, want_status=False)
        ok = ok and dec.checksum() == host_checksum(t.pcm)
        if not ok:
            raise SystemExit(f"bench.py: latency leg {name}: decoded PCM differs from the encoder's input")
        ms = []
        for i in range(3 + args.latency_steps):
            dec.reindex()
            t0 = time.perf_counter()
            dec.decode_all(False, want_status=False)
            dt = (time.perf_counter() - t0) * 1e3
            if i >= 3:
                ms.append((dt, dec.timing()["kernels_ms"]))
        wall = float(np.median([a for a, _ in ms]))
        dev = float(np.median([b for _, b in ms]))
        e = []
        for i in range(2 + max(3, args.latency_steps // 4)):
            t0 = time.perf_counter()
            dec.clear()
            staged.add(dec, Piece(corpus, 0))
            dec.decode_all(host, want_status=False)
            if i >= 2:
                e.append((time.perf_counter() - t0) * 1e3)
        host.free()
    e2e = float(np.median(e))
    return {"workload": corpus.desc, "frames": t.n_frames, "samples": t.n_samples,
            "device_ms": dev, "wall_ms": wall, "value": t.n_samples / (wall * 1e-3) / 1e6,
            "e2e_ms": e2e, "e2e_value": t.n_samples / (e2e * 1e-3) / 1e6, "unit": UNIT,
            # integer-issue accounting of the fused entropy + LPC launch from the committed ncu capture of this config
            "issue": load_profile_json("issue.json").get(name, {}).get("k12_entropy_lpc"),
            "parity": "bit-exact vs the encoder's input + device checksum"}


def copy_ceiling(torch, device, barrier, nbytes=1 << 30, reps=4):
    """What the host can move for this rank while every other rank does the same: one H2D and one D2H stream of
    page-locked 1 GiB copies running at the same time (the shape of the end-to-end pipeline), GB/s each way.
    Under N ranks all of them measure between the same barriers, so the sum is the box's ceiling at N."""
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=device)
    d_out = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def go(n):
        for _ in range(n):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    go(1)
    barrier()
    t0 = time.perf_counter()
    go(reps)
    barrier()
    dt = time.perf_counter() - t0
    return nbytes * reps / dt / 1e9


def load_profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from alac.net_b200 import BatchDecoder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the decode path has no CPU fallback)")
    single = args.single_process and world == 1 and args.gpus > 1
    devices = list(range(args.gpus)) if single else [local]
    torch.cuda.set_device(devices[0])
    from alac.net_b200.shard import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if not (args.no_numa_bind or single) else {"numa_node": None, "cpus": None}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at communicator creation; stdout carries exactly one
        # JSON line, so the banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shards = world if world > 1 else (args.gpus if single else 1)
    name = default_workload(args, shards)
    corpus = make_corpus(name, shards, args)
    pieces = rank_pieces(corpus, world, rank)
    staged = Staged(corpus)
    n_frames = sum(p.f_hi - p.f_lo for p in pieces)
    samples = sum((p.p_hi - p.p_lo) // (corpus.uniq[p.u].cfg.sample_size // 8) for p in pieces)
    pcm_bytes = sum(p.p_hi - p.p_lo for p in pieces)
    comp_bytes = sum(p.b_hi - p.b_lo for p in pieces)

    dec = BatchDecoder(devices=devices, chunk_frames=args.chunk_frames, entropy_lanes=args.entropy_lanes, flags=args.flags)
    t_setup = time.perf_counter()
    for p in pieces:
        staged.add(dec, p)
    total = dec.prepare()              # stage the mdat in HBM + header pre-pass: inputs resident
    t_setup = time.perf_counter() - t_setup

    # ---- parity in the same run, before anything is timed ------------------------------------------
    _, _, _, status = dec.decode_all(False, want_status=True)
    ok = bool((status == 0).all())
    dev_sum = dec.checksum()
    exp_sum = expected_checksum(corpus, dec, pieces)
    ok = ok and dev_sum == exp_sum
    if not ok:
        raise SystemExit(f"bench.py: rank {rank}: decoded PCM does not match the encoder's input "
                         f"(status ok {bool((status == 0).all())}, checksum {dev_sum:#x} vs {exp_sum:#x}) -- refusing to time a wrong decoder")

    # ---- device-resident steps ----------------------------------------------------------------------
    def step():
        dec.reindex()                 # marks the frame index stale: K0 runs again inside the next decode_all
        dec.decode_all(False, want_status=False)
        return dec.timing()

    warm = max(3, args.warmup)
    sampler = ClockSampler(devices[0], period=0.05)
    sampler.start()
    for _ in range(warm):
        step()
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    acc = {"index_ms": 0.0, "entropy_ms": 0.0, "lpc_ms": 0.0, "stereo_ms": 0.0, "kernels_ms": 0.0}
    launches = 0
    chunks = 0
    step_ms, api_ms = [], []
    for _ in range(args.steps):
        ts = time.perf_counter()
        tm = step()
        step_ms.append((time.perf_counter() - ts) * 1e3)
        api_ms.append(tm["total_ms"])
        for k in acc:
            acc[k] += tm[k]
        launches += tm["kernel_launches"]
        chunks = tm["chunks"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.finish()
    dev_ms = acc["kernels_ms"] / args.steps     # CUDA events on the launch streams: first launch -> last kernel done
    wall_ms = wall * 1e3 / args.steps
    # a timed step must not have produced a wrong frame silently: the checksum of the LAST timed step's PCM
    if dec.checksum() != exp_sum:
        raise SystemExit(f"bench.py: rank {rank}: PCM of the last timed step differs from the encoder's input")

    # ---- end to end through the C ABI with host buffers -----------------------------------------------
    e2e = None
    if args.e2e_steps > 0:
        avail = mem_available()
        cap = min(total + 4096, max(1 << 30, int(avail * 0.7) // (world if world > 1 else 1)))   # freed again before the pageable leg
        if args.e2e_buffer_gb > 0:
            cap = min(cap, int(args.e2e_buffer_gb * (1 << 30)))
        ms, ngroups, tm_e2e, ok_e = e2e_leg(dec, staged, pieces, args.e2e_steps, 2, barrier, True, cap)
        if not ok_e:
            raise SystemExit(f"bench.py: rank {rank}: end-to-end PCM differs from the encoder's input")
        e2e = {"ms": ms, "groups": ngroups, "tm": tm_e2e, "cap": cap}
    ceiling = copy_ceiling(torch, torch.device("cuda", devices[0]), barrier) if (e2e and not single) else None
    e2e_pg = None
    if args.e2e_pageable_steps > 0:
        avail = mem_available()
        cap = min(total + 4096, max(1 << 30, int(avail * 0.55) // (world if world > 1 else 1)))
        if args.e2e_buffer_gb > 0:
            cap = min(cap, int(args.e2e_buffer_gb * (1 << 30)))
        staged_pg = Staged(corpus, pinned=False)
        ms, ngroups, _, ok_e = e2e_leg(dec, staged_pg, pieces, args.e2e_pageable_steps, 1, barrier, False, cap)
        if not ok_e:
            raise SystemExit(f"bench.py: rank {rank}: end-to-end (pageable) PCM differs from the encoder's input")
        e2e_pg = {"ms": ms, "groups": ngroups}
    dec.close()

    # ---- max over ranks ---------------------------------------------------------------------------------
    vec = torch.tensor([dev_ms, wall_ms, e2e["ms"] if e2e else 0.0, e2e_pg["ms"] if e2e_pg else 0.0],
                       dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(samples), float(launches), float(n_frames), float(pcm_bytes), float(comp_bytes), float(ceiling or 0.0)],
                       dtype=torch.float64, device="cuda")
    ranges = torch.zeros(world * 2, dtype=torch.float64, device="cuda")
    # this rank's global frame range (for the line's evidence that the shards differ)
    if world > 1:
        first = sum(corpus.track(j).n_frames for j in range(pieces[0].j)) + pieces[0].f_lo if pieces else 0
        ranges[2 * rank], ranges[2 * rank + 1] = float(first), float(first + n_frames)
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(ranges, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, e2e_ms_max, e2e_pg_ms_max = (float(x) for x in vec.tolist())
    samples_all, launches_all, frames_all, pcm_all, comp_all, ceiling_all = (float(x) for x in tot.tolist())

    if rank == 0:
        peak, peak_src = measured_peak()
        stage = {k: acc[k] / args.steps for k in acc}
        b_alg = comp_bytes + pcm_bytes
        frame_lanes = (not (args.flags & 0x42)) and ((args.flags & 0x80) or n_frames // max(1, len(devices)) >= int(os.environ.get("ALACGPU_KF_MIN", 650000)))
        fused = not (args.flags & 2)
        if frame_lanes:
            dom_name, dom_sum = "kf_frames", stage["entropy_ms"] + stage["lpc_ms"]
            dom_what = ("kf_frames<A> + kf_frames<B> (frame-lane entropy + LPC + un-mix/pack, the two phases of one "
                        "kernel template; the class sort in front of phase A is inside the event pair)")
        else:
            key = max(("entropy_ms", "lpc_ms", "stereo_ms"), key=lambda k: stage[k])
            dom_name = {"entropy_ms": "k12_entropy_lpc" if fused else "k1_entropy", "lpc_ms": "k2_lpc",
                        "stereo_ms": "k3_stereo_pack"}[key]
            dom_sum, dom_what = stage[key], dom_name
        stage_sum = stage["index_ms"] + stage["entropy_ms"] + stage["lpc_ms"] + stage["stereo_ms"]
        path_ms = stage["kernels_ms"]
        share = dom_sum / max(stage_sum, 1e-9)
        # chunks run concurrently on several streams, so per-launch event times overlap: the kernel's time is its
        # share of the summed stage times applied to the pipeline's span (first launch -> last kernel done)
        dom_ms = path_ms * share if chunks > 1 else dom_sum
        achieved = b_alg / (dom_ms * 1e-3) / 1e9
        traffic = load_profile_json("traffic.json").get(name, {}).get(dom_name)
        if isinstance(traffic, dict):       # captured on a smaller batch of the same workload: DRAM bytes scale with the frames
            traffic = traffic["bytes"] * (n_frames / max(1, traffic["frames"]))
        issue = load_profile_json("issue.json").get(name, {}).get(dom_name)
        line = {
            "metric": METRIC, "value": samples_all / (wall_ms_max * 1e-3) / 1e6, "unit": UNIT,
            "n_gpus": world if world > 1 else len(devices), "steps": args.steps, "warmup": warm, "ms_per_step": wall_ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {
                "workload": corpus.desc, "scale": args.scale, "tracks": corpus.n_tracks,
                "frames": int(frames_all), "samples": int(samples_all), "compressed_bytes": int(comp_all), "pcm_bytes": int(pcm_all),
                "frames_rank0": n_frames, "compression_ratio": comp_all / max(1.0, pcm_all),
                "l2": "inputs + PCM per step exceed the 126 MB L2 many times over (no flush needed)"
                      if b_alg > 4 * 126e6 else "working set below 4x L2: numbers include L2 hits",
                "parallelism": (f"frame-range shards of ONE batch over {world} ranks (alacgpu_plan_partition / shard.rank_slices), "
                                "no data-path collective" if world > 1 else
                                (f"one alacgpu context over {len(devices)} GPUs (in-library frame-range partition)" if single
                                 else "1 GPU, one decode_all call")),
                "rank_frame_ranges": [[int(ranges[2 * r]), int(ranges[2 * r + 1])] for r in range(world)] if world > 1 else None,
                "parity": "per rank, before timing: device checksum of the decoded PCM == checksum of the encoder's input, every "
                          "frame status OK; after the last timed step the checksum again; end-to-end legs byte-compared",
                "decode_path": "frame lanes (kf_frames)" if frame_lanes else ("fused entropy + LPC (k12)" if fused else "k1 + k2 + k3"),
                "setup_s": t_setup, "numa_bind": numa, "chunks_per_step": chunks,
            },
            "device_ms_per_step": dev_ms_max,
            "step_wall_ms_rank0": [round(x, 2) for x in step_ms], "decode_all_ms_rank0": [round(x, 2) for x in api_ms],
            "stage_ms": stage,
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": dom_what,
                         "scope": "dominant kernel on rank 0: algorithmic bytes of the batch (compressed in + PCM out, "
                                  "intermediates not counted) / its launch time from CUDA events on its launch streams "
                                  "(summed over chunks; chunks overlap on several streams, so the sum is scaled to the "
                                  "pipeline span by the kernel's share); the path is integer-issue bound, see `issue`",
                         "dominant_kernel_ms": dom_ms, "dominant_kernel_ms_summed": dom_sum, "dominant_kernel_share": share,
                         "algorithmic_bytes": b_alg, "issue": issue,
                         "whole_path": {"ms": path_ms, "achieved": b_alg / (path_ms * 1e-3) / 1e9,
                                        "frac": b_alg / (path_ms * 1e-3) / 1e9 / peak}},
        }
        if e2e:
            tm_e = e2e["tm"]
            line["e2e"] = {"value": samples_all / (e2e_ms_max * 1e-3) / 1e6, "unit": UNIT,
                           "h2d_bytes_per_step": int(comp_all + 4 * frames_all), "d2h_bytes_per_step": int(pcm_all),   # whole job
                           "ms_per_step": e2e_ms_max, "steps": args.e2e_steps, "decode_all_calls_per_step": e2e["groups"],
                           "host_buffers": "page-locked (alacgpu_host_alloc)", "host_out_buffer_bytes": e2e["cap"],
                           "achieved_copy_gbs_each_way": (comp_all + pcm_all) / 2 / (e2e_ms_max * 1e-3) / 1e9,
                           "host_copy_ceiling_gbs_each_way": ceiling_all or None,
                           "host_copy_ceiling_note": "sum over ranks of simultaneous page-locked 1 GiB H2D + D2H copies, all ranks between the same barriers",
                           "last_call": {k: tm_e[k] for k in ("h2d_ms", "d2h_ms", "kernels_ms", "total_ms")}}
        if e2e_pg:
            line["e2e_pageable"] = {"value": samples_all / (e2e_pg_ms_max * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_pg_ms_max,
                                    "steps": args.e2e_pageable_steps, "decode_all_calls_per_step": e2e_pg["groups"],
                                    "host_buffers": "plain pageable memory (numpy) for both mdat and PCM"}
        if world == 1 and not single and args.latency_steps > 0 and name == "config4":
            line["latency_legs"] = {k: latency_leg(k, args, local) for k in ("config1", "config2", "config3")}
        if world > 1 and not args.no_multidev_check:
            # the drop-in C# API can only use the in-library form of multi-GPU (ONE alacgpu context over several
            # devices); prove it on this box next to the per-rank numbers: tests/_multidev_case.py decodes a small
            # batch through one context over GPUs 0 and 1 and compares it with the oracle byte for byte
            import subprocess
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_multidev_case.py"), "0", "1"],
                                   capture_output=True, text=True, timeout=600)
                line["config"]["multi_device_context_check"] = ("ok: one context over GPUs 0,1 == oracle" if r.returncode == 0 and
                                                                "multidev case ok" in r.stdout else "FAILED: " + (r.stdout + r.stderr)[-300:])
            except Exception as e:          # noqa: BLE001
                line["config"]["multi_device_context_check"] = f"not run: {e}"
        if not args.no_cpu:
            cores = host_cores()
            _, _, dt0 = cpu_decode_rate(corpus, 8, cores)
            per = int(max(8, min(2048, 12.0 / max(dt0 / 8.0, 1e-6))))
            v, sample, _ = cpu_decode_rate(corpus, per, cores, repeats=2)
            v1, _, _ = cpu_decode_rate(corpus, max(8, per // 4), 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample + "; oracle/ C restatement (C# reference not runnable here)",
                                    "one_core_value": v1}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", help="auto (N=1: config4 = configs[3]; N>1: config5 = configs[4]) | config1..config5")
    ap.add_argument("--scale", type=float, default=1.0, help="duration scale of the tracks (1.0 = BASELINE config)")
    ap.add_argument("--tracks", type=int, default=0, help="tracks of config4 / config5 (0 = 1000 / 1250 per GPU)")
    ap.add_argument("--tracks-per-gpu", type=int, default=0)
    ap.add_argument("--unique", type=int, default=8, help="distinct tracks generated for config4")
    ap.add_argument("--orders", default="0,31", help="workload `fixed`: predictor order range lo,hi of every frame")
    ap.add_argument("--unique-per-kind", type=int, default=1, help="distinct tracks per i mod 10 residue for config5")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-pageable-steps", type=int, default=2)
    ap.add_argument("--e2e-buffer-gb", type=float, default=0.0, help="cap of the host PCM buffer of the e2e legs (0 = what fits)")
    ap.add_argument("--latency-steps", type=int, default=20, help="steps of the configs[0..2] latency legs at N=1 (0 = skip)")
    ap.add_argument("--chunk-frames", type=int, default=0)
    ap.add_argument("--entropy-lanes", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--flags", type=int, default=0,
                    help="ALACGPU_FLAG_* bits: 2 no fusion, 4 no pack fusion, 8 no zero-copy output, 0x40 no frame lanes, 0x80 force frame lanes")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--single-process", action="store_true", help="one process, ONE alacgpu context over --gpus devices")
    ap.add_argument("--no-multidev-check", action="store_true", help="N>1: skip the one-context-over-two-GPUs parity check on rank 0")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
