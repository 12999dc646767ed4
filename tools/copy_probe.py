#!/usr/bin/env python3
"""What the BOX can do: pinned host<->device copies of bench.py's per-step byte counts, no kernels.

    python tools/copy_probe.py                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/copy_probe.py

Every rank copies H2D `--h2d` bytes and D2H `--d2h` bytes per step (defaults: config #2's 155 MB in / 363 MB out) on two
streams, `--steps` times, all ranks at the same time; reported per direction alone and both together (full duplex),
per GPU and for the whole box.  Separates "box ceiling" from "our pipeline" in bench.py's e2e numbers.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h2d", type=int, default=155_000_000)
    ap.add_argument("--d2h", type=int, default=363_000_000)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--bind", action="store_true", help="bind the rank to the NUMA node of its GPU first (bench.py does)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    numa = None
    if args.bind:
        import bench
        numa = bench.bind_near_gpu(local)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h_in = torch.empty(args.h2d, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(args.d2h, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(2)
    d_in = torch.empty(args.h2d, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(args.d2h, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(do_in, do_out):
        def once():
            if do_in:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        for _ in range(3):
            once()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            once()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt
    res = {"n_gpus": world, "h2d_bytes": args.h2d, "d2h_bytes": args.d2h, "steps": args.steps, "numa": numa,
           "host_cpus": os.cpu_count()}
    for name, a, b in (("h2d_alone", True, False), ("d2h_alone", False, True), ("both", True, True)):
        dt = run(a, b)
        res[name] = {"ms_per_step": 1e3 * dt / args.steps,
                     "h2d_gbs_per_gpu": args.h2d * args.steps / dt / 1e9 if a else None,
                     "d2h_gbs_per_gpu": args.d2h * args.steps / dt / 1e9 if b else None,
                     "box_gbs": ((args.h2d if a else 0) + (args.d2h if b else 0)) * args.steps * world / dt / 1e9}
    # what bench.py's e2e would be if it were nothing but these copies: input bytes per step / time of `both`
    res["e2e_ceiling_text_gbs_box"] = world * (args.h2d - 8_000_008) / (res["both"]["ms_per_step"] * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
