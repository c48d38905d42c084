#!/usr/bin/env python
"""bench.py -- decoded PCM Msamples/s of the ALAC frame-decode path on B200.

    python bench.py --gpus N --steps K --warmup W            (ours, N=1)
    torchrun ... bench.py --gpus N --steps K --warmup W      (ours, N>1: one rank per GPU)
    python bench.py --impl reference ...                     (CPU oracle on all host cores)

A "step" is one pass of the whole kernel path (K0 header index + order sort ->
K12 fused entropy + LPC -> K3 stereo/pack) over the workload with the compressed
input already resident in HBM.  `value` = channel values decoded by all ranks / max-over-ranks time.
`e2e` is the same metric through the C ABI with HOST buffers: every step stages
the mdat from pinned host memory (H2D), decodes and copies the PCM back (D2H).

Workload (default): BASELINE.json configs[1] -- synthetic 24-bit stereo 96 kHz
ALAC, 10 min, 4096-sample frames, LPC order 1..31, wasted bytes 0/1/2.  Under
N ranks every rank decodes its own track of that shape (seed + rank): frames
are independent, no collective (scaling: weak).
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


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------
def make_workload(name: str, rank: int, scale: float):
    from tools.alacgen import alacgen as g
    g.build_encoder()
    if name == "config2":
        tracks = [g.track_24_stereo(g.SEED_BASE + 2 + 7919 * rank, 600.0 * scale)]
        desc = "configs[1]: synthetic 24-bit stereo 96 kHz ALAC, 10 min, 4096-sample frames, LPC order 1..31, wasted bytes 0/1/2"
    elif name == "config1":
        tracks = [g.track_16_stereo(g.SEED_BASE + 1 + 7919 * rank, 60.0 * scale)]
        desc = "configs[0]: synthetic 16-bit stereo 44.1 kHz ALAC, 60 s, 4096-sample frames"
    elif name == "config3":
        tracks = [g.track_16_mono_mixed(g.SEED_BASE + 3 + 7919 * rank, 60.0 * scale)]
        desc = "configs[2]: 16-bit mono mixing compressed / uncompressed / Rice-escape frames"
    elif name.startswith("config4"):
        # configs[3]: 1,000 16-bit stereo tracks (~70 h).  `unique` distinct tracks are generated
        # and replicated PHYSICALLY (each replica is staged separately in HBM).
        unique = 8
        base = [g.track_16_stereo(g.SEED_BASE + 1000 + i + 7919 * rank, 252.0 * scale) for i in range(unique)]
        n = int(name.split(":")[1]) if ":" in name else 1000
        tracks = [base[i % unique] for i in range(n)]
        desc = f"configs[3]: batch of {n} synthetic 16-bit stereo tracks x 252 s ({unique} unique, replicated physically)"
    else:
        raise SystemExit(f"unknown workload {name}")
    return tracks, desc


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
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
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

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


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------
def cpu_decode_rate(tracks, budget_frames_per_thread: int, threads: int, repeats: int = 1):
    """One slice of frames per thread (frames are independent, so a thread decodes its own
    contiguous range exactly as one AlacContext would pump it).  -> (Msamples/s, sample text)."""
    from oracle import oracle as o
    o.build()
    t = tracks[0]
    cfg = o.cfg_from(t.cfg)
    nf = t.n_frames
    per = max(1, min(budget_frames_per_thread, nf // max(1, threads)))
    offs = np.concatenate([[0], np.cumsum(t.stsz.astype(np.int64))])
    mdat = np.frombuffer(t.mdat, dtype=np.uint8)
    jobs = []
    for i in range(threads):
        a = (i * per) % max(1, nf - per + 1)
        jobs.append((a, a + per))
    samples = sum(int(t.frame_samples[a:b].sum()) for a, b in jobs) * t.cfg.num_channels

    def work(a, b):
        o.decode_track(cfg, mdat[offs[a]:offs[b]].tobytes(), t.stsz[a:b])

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
    return samples / best / 1e6, f"{threads} threads x {per} frames ({samples} samples) of the workload, best of {repeats}", best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    tracks, desc = make_workload(args.workload, 0, args.scale)
    cores = os.cpu_count() or 1
    # bounded sample: calibrate on a small slice, then size each step so the whole
    # --steps/--warmup run stays near a 90 s budget (at most the full workload per step)
    _, _, dt0 = cpu_decode_rate(tracks, 8, cores)
    per_frame = dt0 / 8.0
    budget = 90.0 / max(1, args.steps + args.warmup)
    per = int(max(8, min(tracks[0].n_frames // cores, budget / max(per_frame, 1e-6))))
    vals = []
    for i in range(args.warmup + args.steps):
        v, sample, dt = cpu_decode_rate(tracks, per, cores)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": desc, "scale": args.scale},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + " per step; oracle/ C restatement of AlacFile.cs (the C# reference cannot run: no .NET in the image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def batch_leg(args, local):
    """Throughput regime (not the headline): a configs[3]-shaped batch -- many 16-bit stereo tracks in ONE
    decode_all -- where there are more frames than lanes and the kernels are issue-bound, not latency-bound."""
    from alac.net_b200 import BatchDecoder, PinnedBuffer
    tracks, desc = make_workload(f"config4:{args.batch_tracks}", 0, 1.0)
    pinned = {}
    for t in tracks:                     # replicas share the pinned source; each is staged separately in HBM
        if id(t) not in pinned:
            pb = PinnedBuffer(len(t.mdat))
            pb.array[:] = np.frombuffer(t.mdat, dtype=np.uint8)
            pinned[id(t)] = pb
    samples = sum(t.n_samples for t in tracks)
    with BatchDecoder(devices=[local], flags=args.flags | (2 if args.no_fusion else 0)) as dec:
        for t in tracks:
            dec.add_track(t.cfg, pinned[id(t)], t.stsz)
        dec.prepare()
        dec.decode_all(False, want_status=False)
        ok = dec.checksum() == sum(host_checksum_at(dec, i, t) for i, t in enumerate(tracks)) % (1 << 64)
        if not ok:
            raise SystemExit("bench.py: batch leg PCM checksum differs from the encoder's input")
        ms = []
        for _ in range(3):
            dec.reindex()
            dec.decode_all(False, want_status=False)
            tm = dec.timing()
            ms.append(tm["kernels_ms"])      # one pipeline: K0 + sort + K12 + K3 (reindex is lazy)
        t_ms = float(np.median(ms))
        stage = {k: tm[k] for k in ("index_ms", "entropy_ms", "lpc_ms", "stereo_ms", "kernels_ms", "chunks")}
    comp = sum(len(t.mdat) for t in tracks)
    pcm = sum(len(t.pcm) for t in tracks)
    return {"workload": desc, "frames": sum(t.n_frames for t in tracks), "samples": samples,
            "device_ms": t_ms, "stage_ms": stage, "value": samples / (t_ms * 1e-3) / 1e6, "unit": UNIT,
            "algorithmic_bytes": comp + pcm, "hbm_gbs": (comp + pcm) / (t_ms * 1e-3) / 1e9,
            "parity": "device checksum of the resident PCM == checksum of the encoder's input"}


def host_checksum_at(dec, i, t):
    from alac.net_b200 import host_checksum
    off, ln = dec.track_pcm_bytes(i)
    return host_checksum(t.pcm, first_word=off // 8)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from alac.net_b200 import BatchDecoder, PinnedBuffer, host_checksum

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the decode path has no CPU fallback)")
    torch.cuda.set_device(local)
    from alac.net_b200.shard import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if not args.no_numa_bind else {"numa_node": None, "cpus": None}
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

    tracks, desc = make_workload(args.workload, rank, args.scale)
    samples = sum(t.n_samples for t in tracks)
    pcm_bytes = sum(len(t.pcm) for t in tracks)
    comp_bytes = sum(len(t.mdat) for t in tracks)
    n_frames = sum(t.n_frames for t in tracks)

    # pinned host copies of the inputs (e2e stages from these every step)
    pinned = []
    for t in tracks:
        pb = PinnedBuffer(len(t.mdat))
        pb.array[:] = np.frombuffer(t.mdat, dtype=np.uint8)
        pinned.append(pb)

    dec = BatchDecoder(devices=[local], chunk_frames=args.chunk_frames, entropy_lanes=args.entropy_lanes,
                       flags=args.flags | (2 if args.no_fusion else 0))
    for t, pb in zip(tracks, pinned):
        dec.add_track(t.cfg, pb, t.stsz)
    total = dec.prepare()              # stage the mdat in HBM + header pre-pass: inputs resident
    host_out = PinnedBuffer(total)

    # ---- parity in the same run: full PCM vs the encoder's input + checksum ----
    out, off, ln, status = dec.decode_all(host_out)
    ok = bool((status == 0).all())
    for t, o_, l_ in zip(tracks, off, ln):
        ok = ok and out[int(o_):int(o_ + l_)].tobytes() == t.pcm
    dec.decode_all(False, want_status=False)      # device-resident copy for the on-device checksum
    dev_sum = dec.checksum()
    ok = ok and dev_sum == host_checksum(out[:total])
    if not ok:
        raise SystemExit("bench.py: decoded PCM does not match the encoder's input -- refusing to time a wrong decoder")

    # ---- device-resident steps ----------------------------------------------------
    def step():
        dec.reindex()                 # marks the frame index stale: K0 runs again inside the next decode_all
        dec.decode_all(False, want_status=False)
        return dec.timing()

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    acc = {"index_ms": 0.0, "entropy_ms": 0.0, "lpc_ms": 0.0, "stereo_ms": 0.0, "kernels_ms": 0.0}
    launches = 0
    for _ in range(args.steps):
        tm = step()
        for k in acc:
            acc[k] += tm[k]
        launches += tm["kernel_launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.finish()
    dev_ms = acc["kernels_ms"] / args.steps     # CUDA events on the launch stream: K0 + sort + K12 + K3 in one pipeline
    wall_ms = wall * 1e3 / args.steps

    # ---- end to end through the C ABI with host buffers -----------------------------
    def e2e_step():
        dec.clear()
        for t, pb in zip(tracks, pinned):
            dec.add_track(t.cfg, pb, t.stsz)
        dec.decode_all(host_out, want_status=False)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
    tm_e2e = dec.timing()

    # ---- max over ranks ----------------------------------------------------------------
    vec = torch.tensor([dev_ms, wall_ms, e2e_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(samples), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, e2e_ms_max = (float(x) for x in vec.tolist())
    samples_all, launches_all = (float(x) for x in tot.tolist())

    if rank == 0:
        peak, peak_src = measured_peak()
        stage = {k: acc[k] / args.steps for k in acc}
        b_alg = comp_bytes + pcm_bytes
        fused = not (args.no_fusion or (args.flags & 2))
        pack_fused = fused and bool(args.flags & 0x20) and not (args.flags & 4)
        names = {"entropy_ms": ("k123_decode (fused entropy + LPC + pack)" if pack_fused else
                                "k12_entropy_lpc (fused entropy + LPC)") if fused else "k1_entropy",
                 "lpc_ms": "k2_lpc", "stereo_ms": "k3_stereo_pack"}
        dom = max(("entropy_ms", "lpc_ms", "stereo_ms"), key=lambda k: stage[k])
        path_ms = stage["kernels_ms"]        # whole pipeline, K0 included (stage["index_ms"] is its K0 part)
        # dominant kernel: algorithmic bytes one launch is responsible for (the whole batch's compressed
        # bytes in + PCM bytes out; intermediates not counted) / its CUDA-event duration
        achieved = b_alg / (stage[dom] * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(args.workload, {}).get(names[dom].split(" ")[0])
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": samples_all / (wall_ms_max * 1e-3) / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": wall_ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {
                "workload": desc, "scale": args.scale, "frames_per_gpu": n_frames,
                "samples_per_gpu": samples, "compressed_bytes_per_gpu": comp_bytes, "pcm_bytes_per_gpu": pcm_bytes,
                "compression_ratio": comp_bytes / max(1, pcm_bytes),
                "l2": "inputs + intermediates + PCM per step exceed the 126 MB L2 (no flush needed)"
                      if b_alg > 2 * 126e6 else "working set below 2x L2: numbers include L2 hits",
                "parallelism": f"frame-range shards, {world} rank(s), no collective",
                "parity": "bit-exact vs encoder input and device checksum, checked in this run",
                "fused_entropy_lpc": fused,
                "numa_bind": numa,
            },
            "device_ms_per_step": dev_ms_max,
            "stage_ms": stage,
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "e2e": {"value": samples_all / (e2e_ms_max * 1e-3) / 1e6, "unit": UNIT,
                    "h2d_bytes_per_step": int(comp_bytes + 4 * n_frames) * world, "d2h_bytes_per_step": int(pcm_bytes) * world,   # whole job
                    "ms_per_step": e2e_ms_max, "steps": args.e2e_steps,
                    "h2d_ms": tm_e2e["h2d_ms"], "d2h_ms": tm_e2e["d2h_ms"], "pipeline_ms": tm_e2e["kernels_ms"],
                    "api_ms": tm_e2e["total_ms"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": names[dom],
                         "scope": "dominant kernel: algorithmic bytes (compressed in + PCM out of the batch, "
                                  "intermediates not counted) / its average launch duration from CUDA events on "
                                  "its launch stream (the event pair also spans the work-list sort and the plane clear in "
                                  "front of it, ~0.1 ms); serial-dependency / ALU-issue bound, see DESIGN.md section 3",
                         "dominant_kernel_ms": stage[dom], "dominant_kernel_share": stage[dom] / max(path_ms, 1e-9),
                         "algorithmic_bytes": b_alg,
                         "whole_path": {"ms": path_ms, "achieved": b_alg / (path_ms * 1e-3) / 1e9,
                                        "frac": b_alg / (path_ms * 1e-3) / 1e9 / peak}},
        }
        if world == 1 and args.batch_tracks > 0:
            line["batch"] = batch_leg(args, local)
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            v, sample, _ = cpu_decode_rate(tracks, max(8, min(2048, tracks[0].n_frames // cores)), cores, repeats=2)
            v1, _, _ = cpu_decode_rate(tracks, 256, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": sample + "; oracle/ C restatement (C# reference not runnable here)",
                                    "one_core_value": v1}
        print(json.dumps(line), flush=True)
    dec.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--scale", type=float, default=1.0, help="duration scale of the workload (1.0 = BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--chunk-frames", type=int, default=0)
    ap.add_argument("--entropy-lanes", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="ALACGPU_FLAG_* bits: 2 no fusion, 4 no pack fusion, 8 no zero-copy output")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--batch-tracks", type=int, default=64, help="tracks of the configs[3]-shaped throughput leg (0 = skip)")
    ap.add_argument("--no-fusion", action="store_true", help="entropy and LPC as two kernels (A/B against the fused launch)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
