#!/usr/bin/env python3
"""bench.py -- throughput of LaTok's tokenization hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tweets|mixed|docs] [--impl reference]

One step = one pass of the hot path (split mask + token spans + CSR offsets) over one batch of
synthetic text.  Default workload: BASELINE config #2, 1 M tweet-sized strings (seed 20240601).

Reported on ONE JSON line (rank 0):
  value        input UTF-8 GB/s over all GPUs, text already resident in HBM (device-pointer C-ABI entry)
  e2e          same metric through the host-buffer C-ABI call: pinned host text -> H2D -> kernels -> D2H results
  roofline     algorithmic bytes (B + C + 8T + 16(S+1)) / CUDA-event time of the tokenize kernel
  cpu_baseline the compiled reference (oracle/_ref) on all host cores over a bounded sample (rank 0, N=1)
N > 1 is launched by torchrun (one rank per GPU); ranks hold different batches (weak scaling), there is
no data-path collective; the span-count exchange is one all_gather of an int64 per rank.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "tweets": dict(desc="config#2: 1M synthetic tweet-sized strings (~140 chars, ASCII-heavy), seed 20240601", n=1_000_000),
    "mixed": dict(desc="config#4: 1M mixed-Unicode strings (~160 chars), seed 20240603", n=1_000_000),
    "docs": dict(desc="config#3: long documents of 64 KB, seed 20240602", n=20_000),
    # strong scaling: ONE batch of ~1e9 characters, cut into byte-balanced ranges of whole strings, one range per rank
    "chars1b": dict(desc="config#5: 1B-character batch of config-#2 text (7.1M strings, seeds 20240605+) sharded by "
                         "byte-balanced string ranges", n=7_100_000),
}
CPU_SAMPLE_STRINGS = {"tweets": 200_000, "mixed": 200_000, "docs": 200, "chars1b": 200_000}


def make_batch(workload: str, n: int, seed_shift: int):
    from latok_b200 import synth
    if workload == "chars1b":      # the whole batch (every rank builds the same one and keeps its own range)
        bufs, offs, base, k = [], [np.zeros(1, dtype=np.int64)], 0, 0
        while n > 0:
            m = min(n, 1_000_000)
            b, o = synth.tweets(m, 20240605 + k)
            bufs.append(b); offs.append(o[1:] + base)
            base += len(b); n -= m; k += 1
        return np.concatenate(bufs), np.concatenate(offs)
    if workload == "tweets":
        return synth.tweets(n, 20240601 + seed_shift)
    if workload == "mixed":
        return synth.mixed_unicode(n, 20240603 + seed_shift)
    return synth.long_docs(n, 65536, 20240602 + seed_shift)


# ------------------------------------------------------------------------------------------ CPU arm
_W = {}


def _cpu_init(kind, shard_strings, shard_packed):
    _W["kind"] = kind
    if kind == "reference":
        from oracle import ref_driver
        ref_driver.ext()
        _W["ref"] = ref_driver
        _W["strings"] = shard_strings
    else:
        from oracle import oracle
        oracle.lib()
        _W["oracle"] = oracle
        _W["packed"] = shard_packed


def _cpu_step(_):
    if _W["kind"] == "reference":
        f = _W["ref"].split_positions   # np.nonzero(gen_split_mask(_gen_parse_matrix(t)))[0]
        n = 0
        for t in _W["strings"]:
            if t:
                n += len(f(t))
        return n
    buf, off = _W["packed"]
    return int(_W["oracle"].tokenize_batch_utf8(buf, off, feats=False)["n_tokens"])


class CpuArm:
    """The reference's CPU path on the host cores: one worker process per core, each holding a
    shard of the sample; a step = every worker runs its shard once."""

    def __init__(self, buf, off, n_sample, cores=None):
        import multiprocessing as mp
        from latok_b200 import synth
        from oracle import ref_driver
        self.kind = "reference" if ref_driver.available() else "port"
        self.cores = cores or os.cpu_count() or 1
        n_sample = min(n_sample, len(off) - 1)
        self.n_strings = n_sample
        self.n_bytes = int(off[n_sample])
        bounds = np.linspace(0, n_sample, self.cores + 1).astype(np.int64)
        ctx = mp.get_context("fork")
        self.pools = []
        for w in range(self.cores):
            a, b = int(bounds[w]), int(bounds[w + 1])
            strings = synth.to_strings(buf, off, a, b) if self.kind == "reference" else None
            packed = (np.ascontiguousarray(buf[off[a]:off[b]]), (off[a:b + 1] - off[a]).copy())
            self.pools.append(ctx.Pool(1, initializer=_cpu_init, initargs=(self.kind, strings, packed)))
        self.sample = (f"first {n_sample} strings ({self.n_bytes / 1e6:.1f} MB) of the workload; call = "
                       + ("np.nonzero(gen_split_mask(_gen_parse_matrix(t)))[0] per string on the compiled reference latok.c"
                          if self.kind == "reference" else "oracle C restatement, batch call, utf-8 in"))

    def step(self):
        t0 = time.perf_counter()
        res = [p.apply_async(_cpu_step, (0,)) for p in self.pools]
        for r in res:
            r.get()
        return time.perf_counter() - t0

    def close(self):
        for p in self.pools:
            p.terminate()


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    wl = args.workload
    n_sample = CPU_SAMPLE_STRINGS[wl]
    buf, off = make_batch(wl, n_sample, 0)
    arm = CpuArm(buf, off, n_sample)
    for _ in range(args.warmup):
        arm.step()
    t = sum(arm.step() for _ in range(args.steps))
    arm.close()
    gbs = arm.n_bytes * args.steps / t / 1e9
    line = {
        "impl": "reference", "metric": "tokenized_text_throughput", "value": gbs, "unit": "GB/s",
        "strings_per_s": arm.n_strings * args.steps / t,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl]["desc"], "step": "bounded sample: " + arm.sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": arm.cores, "kind": arm.kind, "sample": arm.sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.stop_flag, self.ok = [], set(), None, False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ GPU arm
def pinned_array(lib, nbytes, dtype):
    p = C.c_void_p()
    from latok_b200 import _lib
    _lib.check(lib.latok_b200_host_alloc(C.byref(p), max(nbytes, 16)))
    arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(nbytes, 16),))
    return arr[:nbytes].view(dtype), p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tweets", choices=list(WORKLOADS))
    ap.add_argument("--strings", type=int, default=0, help="override strings per GPU -- chars1b: of the whole batch -- (smaller = NOT the named config)")
    ap.add_argument("--resident-batches", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 10)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3
    if world == 1 and args.gpus > 1:
        print(json.dumps({"error": "launch N>1 with torch.distributed.run (one rank per GPU)"}))
        return 2

    import torch
    import torch.distributed as dist
    from latok_b200 import _lib
    from latok_b200.engine import Engine, SPLITS, SPANS, FEATS

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = args.workload
    n_strings = args.strings or WORKLOADS[wl]["n"]
    R = max(1, args.resident_batches)
    classify = wl == "mixed"          # config #4 is quoted with token classification (per-token feature sums) enabled
    what = SPLITS | SPANS | (FEATS if classify else 0)
    lib = _lib.load()

    # ---- synthetic batches (distinct per rank and per resident slot), resident in HBM ----------
    strong = wl == "chars1b"
    if strong:
        from latok_b200.sharding import shard_ranges, slice_shard
        fb, fo = make_batch(wl, n_strings, 0)
        s0, s1 = shard_ranges(fo, world)[rank]
        b, o = slice_shard(fb, fo, s0, s1)
        host = [(np.ascontiguousarray(b), np.ascontiguousarray(o))]
        R = 1                         # one resident batch: a rank's range is far larger than the 126 MB L2
        n_strings = s1 - s0
        del fb, fo
    else:
        host = [make_batch(wl, n_strings, 1000 * rank + r) for r in range(R)]
    dev = [(torch.from_numpy(b.copy()).cuda(), torch.from_numpy(o.copy()).cuda()) for b, o in host]
    eng = Engine(local_rank, max(len(b) for b, _ in host) + 4096, n_strings + 1)

    def run_resident(i):
        b, o = dev[i % R]
        eng.submit_device(b.data_ptr(), o.data_ptr(), o.numel() - 1, b.numel(), what)

    # sizes per batch (also sizes the span buffers so no step re-runs for capacity)
    stats = []
    for i in range(R):
        run_resident(i)
        c, t = eng.sizes()
        stats.append((len(host[i][0]), c, t, len(host[i][1]) - 1))
    for i in range(args.warmup):
        run_resident(i)
    eng.sizes()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count()
    kernel_ms = []
    barrier()
    eng.timer_begin()
    for i in range(args.steps):
        run_resident(i)
    elapsed_ms = eng.timer_end()
    barrier()
    launches = eng.launch_count() - launches0
    # per-launch duration of the dominant kernel, measured live (CUDA events around the tokenize kernel
    # on the stream it is launched on), one step at a time after the timed region
    for i in range(min(args.steps, 50)):
        run_resident(i)
        r = eng.sizes()
        ms = C.c_float(0)
        w = C.c_int64(0)
        _lib.check(lib.latok_b200_last_stats(eng._h, C.byref(ms), C.byref(w)))
        kernel_ms.append(ms.value)
    # keep the GPU under the same load until the clock sampler has enough samples
    t_end = time.time() + 1.0
    while len(sampler.samples) < 40 and time.time() < t_end:
        run_resident(0)
        eng.sizes()
    sampler.stop_flag = True
    sampler.join()

    if world > 1:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        tot = torch.tensor([sum(stats[i % R][0] for i in range(args.steps)),
                            sum(stats[i % R][3] for i in range(args.steps))], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_bytes, total_strings = float(tot[0].item()), float(tot[1].item())
        # the path's only cross-GPU datum: per-rank token counts, to rebase global token offsets
        counts = torch.zeros(world, device="cuda", dtype=torch.int64)
        dist.all_gather_into_tensor(counts, torch.tensor([stats[0][2]], device="cuda", dtype=torch.int64))
    else:
        total_bytes = float(sum(stats[i % R][0] for i in range(args.steps)))
        total_strings = float(sum(stats[i % R][3] for i in range(args.steps)))
    value = total_bytes / (elapsed_ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer C-ABI call (pinned host text in, host arrays out) -----
    # Every step = latok_b200_submit (H2D of that step's text + offsets, kernels) + latok_b200_fetch (D2H of the split
    # mask, spans and both CSR arrays).  Double-buffered: two engines, each driven by its own host thread with its own
    # pinned buffers, so one batch's H2D overlaps the other's D2H (PCIe is full duplex); `single` is one engine alone.
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    hb, ho = host[0]
    B0, C0, T0, S0 = stats[0]

    class E2E:
        def __init__(self, engine):
            self.eng = engine
            self.pin_b, self._p1 = pinned_array(lib, len(hb), np.uint8)
            self.pin_o, self._p2 = pinned_array(lib, 8 * len(ho), np.int64)
            self.pin_b[:] = hb
            self.pin_o[:] = ho
            self.out_splits, self._p3 = pinned_array(lib, C0, np.int8)
            self.out_spans, self._p4 = pinned_array(lib, 8 * T0, np.int32)
            self.out_coff, self._p5 = pinned_array(lib, 8 * (S0 + 1), np.int64)
            self.out_toff, self._p6 = pinned_array(lib, 8 * (S0 + 1), np.int64)
            self.out_feats, self._p7 = pinned_array(lib, 25 * T0 if classify else 16, np.int8)

        def step(self):
            _lib.check(lib.latok_b200_submit(self.eng._h, self.pin_b.ctypes.data, self.pin_o.ctypes.data, S0, what))
            _lib.check(lib.latok_b200_fetch(self.eng._h, self.out_splits.ctypes.data, self.out_coff.ctypes.data,
                                            self.out_spans.ctypes.data, self.out_toff.ctypes.data,
                                            self.out_feats.ctypes.data if classify else None, None))

    def timed(workers, steps_each):
        for w in workers:
            for _ in range(2):
                w.step()
        barrier()
        errs = []

        def loop(w):
            try:
                for _ in range(steps_each):
                    w.step()
            except Exception as exc:      # surfaced below
                errs.append(exc)
        t0 = time.perf_counter()
        if len(workers) == 1:
            loop(workers[0])
        else:
            ths = [threading.Thread(target=loop, args=(w,)) for w in workers]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if errs:
            raise errs[0]
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    w0 = E2E(eng)
    single_s = timed([w0], e2e_steps)
    eng2 = Engine(local_rank, len(hb) + 4096, S0 + 1)
    w1 = E2E(eng2)
    per = max(1, (e2e_steps + 1) // 2)
    e2e_s = timed([w0, w1], per)
    e2e_steps_done = 2 * per
    e2e_value = world * B0 * e2e_steps_done / e2e_s / 1e9
    e2e_single = world * B0 * e2e_steps / single_s / 1e9
    # the second engine's arrays must equal the first's (same batch): cheap end-to-end sanity check of the threaded path
    assert np.array_equal(w0.out_spans, w1.out_spans) and np.array_equal(w0.out_toff, w1.out_toff)
    eng2.close()
    h2d = B0 + 8 * (S0 + 1)
    d2h = C0 + 8 * T0 + 16 * (S0 + 1) + 64 + (25 * T0 if classify else 0)

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg = np.mean([b + c + 8 * t + 16 * (s + 1) + (25 * t if classify else 0) for b, c, t, s in stats])
    k_ms = float(np.mean(kernel_ms)) if kernel_ms else float("nan")
    achieved = alg / (k_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.load(open(tf)).get(wl if not args.strings else "", None)
        except Exception:
            traffic = None

    line = {
        "metric": "tokenized_text_throughput", "value": value, "unit": "GB/s",
        "strings_per_s": total_strings / (elapsed_ms * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl]["desc"] + (f" [OVERRIDE strings={args.strings}]" if args.strings else ""),
                   "strings_per_gpu": n_strings, "bytes_per_step_per_gpu": int(np.mean([s[0] for s in stats])),
                   "chars_per_step_per_gpu": int(np.mean([s[1] for s in stats])),
                   "tokens_per_step_per_gpu": int(np.mean([s[2] for s in stats])),
                   "outputs": "int8 split mask + int32 spans + int64 CSR offsets" + (" + int8[T,25] token feature sums" if classify else ""),
                   "l2_hygiene": f"{R} distinct resident batches rotated; per-step footprint "
                                 f"{alg / 1e6:.0f} MB > 126 MB L2",
                   "sharding": ("one rank per GPU on its byte-balanced range of whole strings of the ONE batch, no data-path "
                                "collective, one all-gather of the per-rank token counts" if strong else
                                "one rank per GPU, independent batches, no data-path collective")},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps_done, "strings_per_s": world * S0 * e2e_steps_done / e2e_s,
                "mode": "double-buffered: 2 engines x 1 host thread per GPU, pinned host buffers, submit + fetch per step",
                "single_engine_value": e2e_single},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "latok::v5::tokenize5_kernel<kFeats>" if classify else "latok::v5::tokenize5_kernel", "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": float(alg), "peak_source": peak_src,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "clocks": sampler.summary(),
    }

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            arm = CpuArm(hb, ho, CPU_SAMPLE_STRINGS[wl])
            arm.step()
            t = sum(arm.step() for _ in range(3))
            arm.close()
            line["cpu_baseline"] = {"value": arm.n_bytes * 3 / t / 1e9, "unit": "GB/s", "cores": arm.cores,
                                    "kind": arm.kind, "sample": arm.sample,
                                    "strings_per_s": arm.n_strings * 3 / t}
        except Exception as exc:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "unavailable", "sample": repr(exc)}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
