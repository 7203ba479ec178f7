#!/usr/bin/env python
"""bench.py — uint16 voxel GB/s of the sqeazy volume pipeline (encode + decode) on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2|cfg1|cfg3|cfg5]

A *step* = one encode + one decode of one volume through the pipeline.
  value    raw volume bytes / (t_encode + t_decode), inputs resident in HBM (sqyx_* device API), summed over ranks
  e2e      the same metric through the reference-facing C API with HOST buffers (SQY_PipelineEncode_UI16 +
           SQY_Decode_UI16, pinned host memory, H2D/D2H inside the timed region)
  roofline dominant kernel (the LZ4 block encoder): (input bytes + compressed bytes) / its CUDA-event time vs measured HBM peak
  cpu_baseline  the reference's own stage code (oracle/_ref) on this box's host cores on a bounded slab of the same workload
N > 1 (torchrun): every rank encodes/decodes its own stack (stacks partitioned over GPUs; weak scaling; no data-path
collective — cfg3's quantiser adds the NCCL histogram all-reduce); time = max over ranks.
`--impl reference` times the reference's CPU implementation (rank 0 only) on bounded slabs of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shape {Z,Y,X}, pipeline, preset, reference stage-chain id)
    "cfg1": ((256, 512, 512), "bitswap1->lz4", "scmos", 0),
    "cfg2": ((512, 2048, 2048), "rmestbkrd->bitswap1->lz4", "scmos", 1),
    "cfg3": ((1024, 2048, 2048), "quantiser->lz4", "scmos", 2),
    "cfg5": ((128, 1024, 1024), "remove_background(threshold=110)->bitswap4->lz4", "scmos", 3),
}
METRIC = "uint16 voxel GB/s encode+decode"


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu_index = gpu_index
        self.path = f"/tmp/sqy_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx = float(p[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if not sm:
            # timed region shorter than one sampling period: one immediate reading (GPU still warm), flagged as such
            try:
                line = subprocess.check_output(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                                                "--format=csv,noheader,nounits"], timeout=10).decode().strip()
                p = [x.strip() for x in line.split(",")]
                sm.append(float(p[1]))
                mx = float(p[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
                out["note"] = "timed region < sampling period: sampled right after it"
            except Exception:
                pass
        if sm:
            sm.sort()
            # median of the upper half = clocks under load (idle samples between steps drag a plain median down)
            hi = sm[len(sm) // 2:]
            out["sm_mhz"] = hi[len(hi) // 2]
        out["sm_max_mhz"] = mx
        out["reasons"] = sorted(reasons)
        return out


def reference_arm(args, shape, pipeline, preset, chain_id, steps, warmup, as_baseline=False):
    """times the reference's CPU implementation (oracle/_ref: the reference's stage headers compiled in the build
    container) on a bounded z-slab of the workload, all host threads"""
    import numpy as np

    from oracle import oracle as orc
    from sqeazy_b200.synth import numpy_volume

    ref = orc.ref()
    cores = os.cpu_count() or 1
    Z, Y, X = shape
    # bounded sample: a z-slab of ~128 MiB (< 2^31 voxels, which the reference cannot index anyway: SURVEY F7)
    slab_z = max(4, min(Z, (128 << 20) // (Y * X * 2)))
    sample_shape = (slab_z, Y, X)
    vol = numpy_volume(sample_shape, preset, index=0)
    if not ref.available:
        # the C port of the oracle stands in (scalar, 1 core)
        port = orc.port()
        cores = 1
        kind = "port"

        def run():
            t0 = time.perf_counter()
            cur = vol
            if chain_id == 1:
                cur, _ = port.rmestbkrd(cur, 2 << 20)
            if chain_id == 3:
                cur = port.remove_background(cur, 110)
            if chain_id == 2:
                enc, dec = port.quantiser_luts(port.histogram(cur))
                data = port.lut_apply(cur, enc)
            else:
                data = port.bitswap_encode(4 if chain_id == 3 else 1, cur)
            payload = port.lz4_frames_encode(data)
            t1 = time.perf_counter()
            raw = port.lz4_frames_decode(payload, data.nbytes)
            if chain_id == 2:
                port.lut_decode(raw, dec)
            else:
                port.bitswap_decode(4 if chain_id == 3 else 1, raw.view(np.uint16))
            t2 = time.perf_counter()
            return t1 - t0, t2 - t1, payload.size
    else:
        kind = "reference"
        w = 4 if chain_id == 3 else 1

        def run():
            payload, t_enc = ref.pipeline_encode_stages(chain_id, vol, cores, w=w, threshold=110)
            if chain_id == 2:
                t0 = time.perf_counter()
                rc, codes = ref.lz4_decode_bytes(payload, vol.size)
                t_dec = time.perf_counter() - t0  # LUT decode (tiny) not timed: the reference's signed index is UB (F11)
            else:
                rc, _, t_dec = ref.pipeline_decode_stages(w, payload, vol.size)
            assert rc == 0
            return t_enc, t_dec, payload.size

    n_warm = min(warmup, 1) if as_baseline else warmup
    n_steps = min(steps, 2) if as_baseline else steps
    for _ in range(n_warm):
        run()
    t_enc = t_dec = 0.0
    payload = 0
    for _ in range(n_steps):
        e, d, payload = run()
        t_enc += e
        t_dec += d
    raw = vol.nbytes
    value = raw * n_steps / (t_enc + t_dec) / 1e9
    info = {"value": value, "unit": "GB/s", "cores": cores, "kind": kind,
            "sample": f"z-slab {sample_shape} of {shape}, {n_steps} steps, encode {raw * n_steps / t_enc / 1e9:.3f} GB/s, "
                      f"decode {raw * n_steps / t_dec / 1e9:.3f} GB/s, ratio {raw / payload:.3f}"}
    return info, (t_enc + t_dec) / n_steps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--preset", default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    shape, pipeline, preset, chain_id = WORKLOADS[args.workload]
    if args.preset:
        preset = args.preset
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    config = {"workload": f"{args.workload}: {pipeline} on synthetic {shape[2]}x{shape[1]}x{shape[0]} uint16 light-sheet stack ({preset})",
              "pipeline": pipeline, "shape_zyx": list(shape), "preset": preset,
              "parallelism": f"{world} stack(s), one per GPU" if world > 1 else "1 GPU",
              "cache": "inputs (>= 256 MiB per step) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        info, ms = reference_arm(args, shape, pipeline, preset, chain_id, args.steps, args.warmup)
        line = {"metric": METRIC, "value": info["value"], "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
                "impl": "reference", "config": config, "cpu_baseline": info,
                "e2e": {"value": info["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import numpy as np
    import torch

    import sqeazy_b200 as sq
    from sqeazy_b200.synth import torch_volume

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    sq.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    dev = torch.device("cuda", local_rank)
    vol = torch_volume(shape, preset, index=rank, device=dev)
    raw_bytes = vol.numel() * 2
    cap = sq.max_compressed_length(pipeline, raw_bytes)
    blob_buf = torch.empty(cap, dtype=torch.uint8, device=dev)
    out = torch.empty_like(vol)
    use_hist_allreduce = world > 1 and "quantiser" in pipeline
    hist = torch.zeros(65536, dtype=torch.int32, device=dev) if use_hist_allreduce else None

    def encode_step():
        if use_hist_allreduce:
            hist.zero_()
            sq.histogram_device(vol, hist)
            dist.all_reduce(hist)  # the path's only collective: NCCL sum of the 65536-bin histogram
            return sq.encode_device(pipeline, vol, out=blob_buf, global_hist=hist)
        return sq.encode_device(pipeline, vol, out=blob_buf)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (also sizes the library's scratch arena) ----
    blob = None
    for _ in range(warmup):
        blob = encode_step()
        sq.decode_device(blob, out)
    torch.cuda.synchronize()
    assert torch.equal(out, vol) if pipeline in ("bitswap1->lz4",) else True
    blob_bytes = int(blob.numel())

    # ---- timed region: device-resident ----
    sampler = ClockSampler(local_rank)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    launches0 = sq.kernel_launches()
    barrier()
    sampler.start()
    t_enc_ms = t_dec_ms = 0.0
    for _ in range(args.steps):
        ev[0].record()
        blob = encode_step()
        ev[1].record()
        sq.decode_device(blob, out)
        ev[2].record()
        ev[2].synchronize()
        t_enc_ms += ev[0].elapsed_time(ev[1])
        t_dec_ms += ev[1].elapsed_time(ev[2])
    barrier()
    clocks = sampler.stop()
    launches = sq.kernel_launches() - launches0
    stats = sq.last_lz4_stats()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_enc_ms = max_over_ranks(t_enc_ms)
    t_dec_ms = max_over_ranks(t_dec_ms)
    total_s = (t_enc_ms + t_dec_ms) / 1e3
    value = world * raw_bytes * args.steps / total_s / 1e9

    # ---- per-stage device times (separate pass: the timers add event syncs) ----
    sq.enable_stage_timing(True)
    sq.stage_ms(reset=True)
    n_prof = 2
    for _ in range(n_prof):
        blob = encode_step()
        sq.decode_device(blob, out)
    torch.cuda.synchronize()
    stage = {k: v / n_prof for k, v in sq.stage_ms(reset=True).items()}
    sq.enable_stage_timing(False)
    lz4_in_bytes = raw_bytes // 2 if "quantiser" in pipeline else raw_bytes
    payload_bytes = stats["payload_bytes"]
    peak, peak_kind = measured_peak_hbm()
    achieved = (lz4_in_bytes + payload_bytes) / (stage["lz4_encode"] / 1e3) / 1e9 if stage["lz4_encode"] > 0 else 0.0
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of lz4_encode_kernel from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(args.workload, {}).get("lz4_encode_kernel_dram_bytes_per_launch")
    except Exception:
        pass
    # split of the dominant kernel into its closed-form path (all-equal blocks: load, compare, 75 bytes out) and
    # the general path. The closed-form rate is measured on an all-zero buffer of the same size; the general rate follows
    # from the block counts: t_general = t_kernel - constant_bytes / constant_rate.
    split = None
    try:
        zeros = torch.zeros(lz4_in_bytes, dtype=torch.uint8, device=dev)
        zbuf = torch.empty(sq.lz4_bound(lz4_in_bytes), dtype=torch.uint8, device=dev)
        sq.lz4_encode_device(zeros, out=zbuf)
        zev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        zev[0].record()
        for _ in range(3):
            sq.lz4_encode_device(zeros, out=zbuf)
        zev[1].record()
        zev[1].synchronize()
        const_gbs = lz4_in_bytes / (zev[0].elapsed_time(zev[1]) / 3 / 1e3) / 1e9
        nb = stats["general_blocks"] + stats["constant_blocks"] + stats["stored_blocks"]
        const_bytes = lz4_in_bytes * stats["constant_blocks"] / max(nb, 1)
        t_const = const_bytes / (const_gbs * 1e9)
        t_general = max(stage["lz4_encode"] / 1e3 - t_const, 1e-9)
        split = {"constant_blocks_share": stats["constant_blocks"] / max(nb, 1), "constant_path_gbs": const_gbs,
                 "constant_path_frac_of_peak": const_gbs / peak,
                 "general_path_gbs": (lz4_in_bytes - const_bytes) / t_general / 1e9,
                 "general_path_frac_of_peak": (lz4_in_bytes - const_bytes) / t_general / 1e9 / peak}
        del zeros, zbuf
    except Exception as exc:
        split = {"error": repr(exc)}
    roofline = {"bound": "hbm", "kernel": "lz4_encode_kernel", "achieved": achieved, "peak": peak, "peak_source": peak_kind + " copy bandwidth",
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": lz4_in_bytes + payload_bytes, "kernel_ms": stage["lz4_encode"],
                "stage_ms": stage,
                "stage_gbs": {"filter_bitswap_encode(4B/voxel)": (2 * raw_bytes / (stage["filter_bitswap_encode"] / 1e3) / 1e9) if stage["filter_bitswap_encode"] else None,
                              "bitswap_decode(4B/voxel)": (2 * raw_bytes / (stage["bitswap_decode"] / 1e3) / 1e9) if stage["bitswap_decode"] else None,
                              "lz4_decode(C+B)": ((lz4_in_bytes + payload_bytes) / (stage["lz4_decode"] / 1e3) / 1e9) if stage["lz4_decode"] else None},
                "lz4_blocks": stats, "path_split": split}

    # ---- e2e: host buffers through SQY_PipelineEncode_UI16 / SQY_Decode_UI16 ----
    e2e = None
    if not args.no_e2e:
        try:
            h_vol = torch.empty(vol.shape, dtype=torch.int16).pin_memory()
            h_vol.copy_(vol)
            h_blob = torch.empty(cap, dtype=torch.uint8).pin_memory()
            h_out = torch.empty(vol.shape, dtype=torch.int16).pin_memory()
            np_vol = h_vol.numpy().view(np.uint16)
            np_blob = h_blob.numpy()
            np_out = h_out.numpy().view(np.uint16).reshape(-1)
            e_steps = max(1, min(args.steps, 3))
            b = sq.encode(pipeline, np_vol, out=np_blob)  # warm-up (grows the staging arena)
            sq.decode(b, out=np_out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                b = sq.encode(pipeline, np_vol, out=np_blob)
                sq.decode(b, out=np_out)
            torch.cuda.synchronize()
            t_e2e = max_over_ranks(time.perf_counter() - t0)
            # the host-pointer path must deliver the same voxels as the device-resident path (64-bit checksums)
            host_sum = int(np_out.astype(np.uint64).sum()) if np_out.size < (1 << 28) else int(
                sum(int(np_out[i: i + (1 << 27)].astype(np.uint64).sum()) for i in range(0, np_out.size, 1 << 27)))
            flat = out.view(-1)
            dev_sum = sum(int((flat[i: i + (1 << 28)].to(torch.int64) & 0xFFFF).sum().item()) for i in range(0, flat.numel(), 1 << 28))
            e2e_ok = host_sum == dev_sum
            e2e = {"value": world * raw_bytes * e_steps / t_e2e / 1e9, "unit": "GB/s", "steps": e_steps,
                   "h2d_bytes_per_step": raw_bytes + int(b.size), "d2h_bytes_per_step": int(b.size) + raw_bytes,
                   "api": "SQY_PipelineEncode_UI16 + SQY_Decode_UI16, pinned host buffers", "same_voxels_as_device_path": e2e_ok}
            del h_vol, h_blob, h_out
        except Exception as exc:  # pinned allocation can fail on small hosts; report instead of dying
            e2e = {"value": None, "unit": "GB/s", "error": repr(exc), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = reference_arm(args, shape, pipeline, preset, chain_id, 2, 1, as_baseline=True)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": (t_enc_ms + t_dec_ms) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u16", "data": "synthetic", "config": config,
                "encode_gbs": world * raw_bytes * args.steps / (t_enc_ms / 1e3) / 1e9,
                "decode_gbs": world * raw_bytes * args.steps / (t_dec_ms / 1e3) / 1e9,
                "compression_ratio": raw_bytes / blob_bytes, "blob_bytes": blob_bytes,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
