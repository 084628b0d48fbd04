#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native VietASR CTC hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode f16x3|f16x1|fp32] [--impl reference]

A "step" is one pass of the whole hot path (log-mel front end -> QuartzNet encoder -> CTC decoder
+ greedy argmax -> CTC collapse) over one batch of synthetic 16 kHz clips.  Workload at every N:
BASELINE.json configs[2] per GPU - QuartzNet15x5, batch 256 x 5 s clips (weak scaling: N GPUs process
N*256 clips; configs[3]'s 8-GPU run is the same shape with its 1024 clips split 128 per GPU).

One JSON line is printed by rank 0 (see the contract in the task statement):
  value      audio-seconds/s, whole job, inputs already resident in HBM, CUDA-event time, max over ranks
  e2e        same metric through the host-buffer C-ABI call (H2D of the waveforms and D2H of the collapsed
             ids inside the timed region; at N>1: per-rank pinned host shard -> H2D -> compute -> NCCL gather of the
             ids to rank 0 -> D2H; --e2e-scatter: rank-0 H2D -> NCCL scatter -> compute -> NCCL gather -> D2H)
  roofline   dominant kernel family = the fused sub-block kernel (78 launches/step for 15x5), timed live with
             CUDA events around the encoder stage inside the timed region
  cpu_baseline  the oracle (a port of the reference's torch-CPU arithmetic) on this box's host cores,
             bounded sample, rank 0 only
`--impl reference` times that CPU port alone with all host threads and prints the same line shape.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "quartznet15x5"
B_PER_GPU = 256
CLIP_S = 5.0
SR = 16000
L = int(CLIP_S * SR)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "bf16_tflops_burst": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "source": "fallback"}


def load_traffic(mode):
    """dram__bytes_read.sum + dram__bytes_write.sum of the encoder kernels of ONE step from the committed ncu
    capture (profiles/r1_traffic.json, f16x3, same workload); None for other modes."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if mode != "f16x3" or not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d.get("traffic_bytes_per_step", d.get("traffic_bytes_per_launch", 0) * d.get("launches", 0))


def algorithmic_counts(jasper, feat_in, T_f, num_classes):
    """Fused-ideal fp32 bytes and flops per utterance: each sub-block reads its
    input once and writes its output once, the residual branch re-reads the block input once."""
    T = T_f
    cin = feat_in
    bytes_act = 0
    fl_dw = fl_pw = fl_res = 0
    n_sub = 0
    wbytes = 0
    layers = []     # (bytes, dw flops, 1x1 flops) per sub-block and utterance: for the per-layer composite floor
    for blk in jasper:
        k, s, d = blk["kernel"][0], blk["stride"][0], blk["dilation"][0]
        pad = (d * k) // 2 - 1 if d > 1 else k // 2
        cout = blk["filters"]
        block_cin, block_T = cin, T
        c = cin
        for r in range(blk["repeat"]):
            T_out = (T + 2 * pad - d * (k - 1) - 1) // s + 1
            bytes_act += 4 * (c * T + cout * T_out)
            if blk.get("separable", False):
                fl_dw += 2 * c * k * T_out
                wbytes += 4 * c * k
            fl_pw += 2 * c * cout * T_out
            wbytes += 4 * c * cout
            n_sub += 1
            layers.append([4 * (c * T + cout * T_out), 2 * c * k * T_out if blk.get("separable", False) else 0,
                           2 * c * cout * T_out])
            c, T = cout, T_out
        if blk["residual"]:
            bytes_act += 4 * block_cin * block_T
            fl_res += 2 * block_cin * cout * T
            wbytes += 4 * block_cin * cout
            layers[-1][0] += 4 * block_cin * block_T
            layers[-1][2] += 2 * block_cin * cout * T
        cin = cout
    fl_dec = 2 * cin * num_classes * T
    return {"bytes_act": bytes_act, "flops_dw": fl_dw, "flops_pw": fl_pw, "flops_res": fl_res,
            "flops_dec": fl_dec, "n_sub": n_sub, "weight_bytes": wbytes, "T_e": T, "layers": layers}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.path = f"/tmp/vasr_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def synth_batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    wave = (0.1 * torch.randn(B, L, generator=g)).clamp_(-1.0, 1.0)
    length = torch.full((B,), L, dtype=torch.int64)
    return wave, length


def load_weights_into(eng, V):
    wdir = os.path.join(ROOT, "weights", "en15x5")
    if os.path.exists(os.path.join(wdir, "JasperEncoder.pt")):
        eng.encoder.restore_from(os.path.join(wdir, "JasperEncoder.pt"))
        eng.decoder.restore_from(os.path.join(wdir, "JasperDecoderForCTC.pt"))
        return "shipped QuartzNet15x5 checkpoint"
    return "random-init (xavier) weights of the QuartzNet15x5 architecture"


def pick_threads(md):
    """The reference path is many small convolutions: torch's intra-op pool scales poorly past a few
    dozen threads.  Probe a short pass at several pool sizes and keep the fastest (all cores are
    available to it; `cores` in the JSON is the pool size actually used)."""
    from oracle import quartznet_oracle as O
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu})
    jasper = md["JasperEncoder"]["jasper"]
    enc_sd, dec_sd = O.random_state_dicts(jasper, 64, len(md["labels"]), seed=1)
    wave, length = synth_batch(4, 99)
    best, best_t = None, None
    for t in cands:
        torch.set_num_threads(t)
        O.full_path(enc_sd, dec_sd, jasper, wave[:1], length[:1])
        t0 = time.perf_counter()
        O.full_path(enc_sd, dec_sd, jasper, wave, length)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, best_t = dt, t
        if dt > 20:
            break
    return best_t


def cpu_port_time(md, B_cpu, iters, threads):
    """Oracle (port of the reference's torch-CPU path) on the host cores: audio-seconds/s."""
    from oracle import quartznet_oracle as O
    torch.set_num_threads(threads)
    jasper = md["JasperEncoder"]["jasper"]
    wdir = os.path.join(ROOT, "weights", "en15x5")
    if os.path.exists(os.path.join(wdir, "JasperEncoder.pt")):
        enc_sd = torch.load(os.path.join(wdir, "JasperEncoder.pt"), map_location="cpu")
        dec_sd = torch.load(os.path.join(wdir, "JasperDecoderForCTC.pt"), map_location="cpu")
    else:
        enc_sd, dec_sd = O.random_state_dicts(jasper, 64, len(md["labels"]), seed=1)
    wave, length = synth_batch(B_cpu, 4321)
    blank = len(md["labels"])

    def one():
        r = O.full_path(enc_sd, dec_sd, jasper, wave, length)
        return O.ctc_collapse(r["ids"].numpy(), blank)

    one()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter(); one(); ts.append(time.perf_counter() - t0)
    return B_cpu * CLIP_S / statistics.median(ts), statistics.median(ts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "viet-asr_b200"))
    import configs as cfgs      # plain module import: the reference arm never loads the CUDA library
    md = cfgs.quartznet15x5()
    threads = pick_threads(md)
    B_cpu = 16
    steps, warm = max(1, args.steps), max(0, args.warmup)
    from oracle import quartznet_oracle as O
    torch.set_num_threads(threads)
    jasper = md["JasperEncoder"]["jasper"]
    wdir = os.path.join(ROOT, "weights", "en15x5")
    if os.path.exists(os.path.join(wdir, "JasperEncoder.pt")):
        enc_sd = torch.load(os.path.join(wdir, "JasperEncoder.pt"), map_location="cpu")
        dec_sd = torch.load(os.path.join(wdir, "JasperDecoderForCTC.pt"), map_location="cpu")
    else:
        enc_sd, dec_sd = O.random_state_dicts(jasper, 64, len(md["labels"]), seed=1)
    wave, length = synth_batch(B_cpu, 4321)
    blank = len(md["labels"])
    budget_s = 150.0
    t_start = time.perf_counter()
    ts = []
    for i in range(warm + steps):
        t0 = time.perf_counter()
        r = O.full_path(enc_sd, dec_sd, jasper, wave, length)
        O.ctc_collapse(r["ids"].numpy(), blank)
        dt = time.perf_counter() - t0
        if i >= warm:
            ts.append(dt)
        if time.perf_counter() - t_start > budget_s and len(ts) >= 1:
            break
    med = statistics.median(ts)
    v = B_cpu * CLIP_S / med
    sample = f"{B_cpu} x {CLIP_S:.0f} s synthetic clips per step ({len(ts)} timed steps), torch CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": "audio-seconds/sec (RTF^-1) QuartzNet15x5 16kHz", "value": v,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": len(ts), "warmup": warm, "ms_per_step": med * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"QuartzNet15x5 greedy CTC, {B_PER_GPU} x 5 s synthetic 16 kHz clips per GPU "
                               f"(CPU arm: bounded sample of {B_cpu} clips per step)"},
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads, "host_cpus": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default=os.environ.get("VASR_GEMM_MODE", "f16x3"), choices=["fp32", "f16x3", "f16x1"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-scatter", action="store_true", help="N>1: rank-0 H2D of the whole batch + NCCL scatter instead of per-rank host shards")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import viet_asr_b200 as V
    from viet_asr_b200 import dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    B = args.batch

    md = V.configs.MODELS[MODEL]()
    jasper = md["JasperEncoder"]["jasper"]
    V.NeuralModuleFactory(placement=V.DeviceType.GPU)
    eng = V.VietASR(model_definition=md, gemm_mode=args.mode)
    weights_note = load_weights_into(eng, V)
    blank = len(md["labels"])

    wave_h, len_h = synth_batch(B, 1234 + rank)
    wave_d, len_d = wave_h.to(dev), len_h.to(dev)
    T_f = eng.preprocessor.num_frames(L)
    counts = algorithmic_counts(jasper, 64, T_f, blank + 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    # ------------------------------------------------------------ leg 1: inputs resident in HBM
    def step_device(marks=None):
        feat, seq = eng.preprocessor.forward_channels_last(wave_d, len_d)
        if marks: marks[0].record()
        enc, enc_len = eng.encoder.forward_channels_last(feat, seq)
        if marks: marks[1].record()
        _, ids = eng.decoder.forward_channels_last(enc, False)
        out_ids, out_len = V.ctc_collapse(ids, blank)
        return out_ids, out_len

    for _ in range(W):
        step_device()
    barrier()
    _feat, _seq = eng.preprocessor.forward_channels_last(wave_d, len_d)
    _l0 = V._lib.launch_count()
    eng.encoder.forward_channels_last(_feat, _seq)
    enc_launches = int(V._lib.launch_count() - _l0)      # kernels of one encoder pass (incl. the length kernel)
    del _feat, _seq
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = V._lib.launch_count()
    step_ev = [(ev(), ev(), ev(), ev()) for _ in range(K)]
    barrier()
    for i in range(K):
        s, e0, e1, e = step_ev[i]
        s.record()
        step_device((e0, e1))
        e.record()
    barrier()
    launches = V._lib.launch_count() - launches0
    total_ms = step_ev[0][0].elapsed_time(step_ev[-1][3])
    step_ms = [a.elapsed_time(d) for a, _, _, d in step_ev]
    enc_ms = [b.elapsed_time(c) for _, b, c, _ in step_ev]
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / K
    value = world * B * CLIP_S / (ms_per_step * 1e-3)
    enc_ms_avg = max_over_ranks(sum(enc_ms) / K)

    # ------------------------------------------------------------ leg 2: end to end, host buffers
    T_e = eng.out_frames(L)
    if world == 1:
        wave_p, len_p = wave_h.pin_memory(), len_h.pin_memory()
        oid_p = torch.empty((B, T_e), dtype=torch.int32).pin_memory()
        oln_p = torch.empty((B,), dtype=torch.int32).pin_memory()

        def step_e2e():
            eng.transcribe_host_ids(wave_p, len_p, oid_p, oln_p)
        h2d = B * L * 4 + B * 8
        d2h = B * T_e * 4 + B * 4
    else:
        # N > 1: one process per GPU, each with its own pinned host shard (the batch is sharded on the host side, so the
        # waveforms cross each GPU's own PCIe link in parallel); NCCL carries the result gather to rank 0, whose D2H of
        # the gathered ids closes the step.  `--e2e-scatter` measures the rank-0-H2D + NCCL-scatter variant instead.
        GB = world * B
        wave_p, len_p = wave_h.pin_memory(), len_h.pin_memory()
        w_dev = torch.empty((B, L), dtype=torch.float32, device=dev)
        l_dev = torch.empty((B,), dtype=torch.int64, device=dev)
        if rank == 0:
            oid_p = torch.empty((GB, T_e), dtype=torch.int32).pin_memory()
            oln_p = torch.empty((GB,), dtype=torch.int32).pin_memory()
        if args.e2e_scatter and rank == 0:
            gw = torch.cat([synth_batch(B, 1234 + r)[0] for r in range(world)]).pin_memory()
            gl = torch.full((GB,), L, dtype=torch.int64).pin_memory()
            gw_d = torch.empty((GB, L), dtype=torch.float32, device=dev)
            gl_d = torch.empty((GB,), dtype=torch.int64, device=dev)
        else:
            gw_d = gl_d = None

        def step_e2e():
            if args.e2e_scatter:
                if rank == 0:
                    gw_d.copy_(gw, non_blocking=True); gl_d.copy_(gl, non_blocking=True)
                w, ln = D.scatter_batch(gw_d, gl_d, GB, L, dev)
            else:
                w_dev.copy_(wave_p, non_blocking=True); l_dev.copy_(len_p, non_blocking=True)
                w, ln = w_dev, l_dev
            r = eng.forward_device(w, ln)
            gi, gn = D.gather_results(r["out_ids"], r["out_len"], GB)
            if rank == 0:
                oid_p.copy_(gi, non_blocking=True); oln_p.copy_(gn, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d = GB * L * 4 + GB * 8
        d2h = GB * T_e * 4 + GB * 4
    for _ in range(W):
        step_e2e()
    barrier()
    s2, e2 = ev(), ev()
    s2.record()
    for _ in range(K):
        step_e2e()
    e2.record()
    barrier()
    e2e_ms = max_over_ranks(s2.elapsed_time(e2)) / K
    e2e_value = world * B * CLIP_S / (e2e_ms * 1e-3)
    clocks = sampler.stop() if sampler else None

    # ------------------------------------------------------------ single-utterance latency (the web-app case)
    lat = None
    if rank == 0:
        w1, l1 = synth_batch(1, 777)
        w1p, l1p = w1.pin_memory(), l1.pin_memory()
        o1 = torch.empty((1, T_e), dtype=torch.int32).pin_memory(); n1 = torch.empty((1,), dtype=torch.int32).pin_memory()
        for _ in range(5):
            eng.transcribe_host_ids(w1p, l1p, o1, n1)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); eng.transcribe_host_ids(w1p, l1p, o1, n1); ts.append((time.perf_counter() - t0) * 1e3)
        lat = {"batch": 1, "clip_seconds": CLIP_S, "p50_ms": statistics.median(ts), "p90_ms": sorted(ts)[26],
               "what": "vasr_transcribe_host wall time (H2D + whole path + D2H + sync), greedy"}

    # sanity: the device leg and the host leg agree on the transcript of this rank's batch (world==1)
    legs_agree = None
    if world == 1:
        o_ids, o_len = step_device()
        legs_agree = bool(torch.equal(o_ids.cpu(), oid_p) and torch.equal(o_len.cpu(), oln_p))
        if not legs_agree:
            print("WARNING: device-resident and host-buffer legs produced different transcripts", file=sys.stderr)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    enc_bytes = B * counts["bytes_act"] + counts["weight_bytes"]
    enc_flops_tc = B * (counts["flops_pw"] + counts["flops_res"])
    enc_flops_dw = B * counts["flops_dw"]
    n_sub = counts["n_sub"]
    ach_gbs = enc_bytes / (enc_ms_avg * 1e-3) / 1e9
    ach_tf = enc_flops_tc / (enc_ms_avg * 1e-3) / 1e12
    ach_dw_tf = enc_flops_dw / (enc_ms_avg * 1e-3) / 1e12
    # FP32 pipe peak from the committed micro-benchmark (tools/ubench/fma_rate, profiles/r1_ubench_fma_rate.log):
    # packed FFMA2 sustains 125 FMA lanes per clock per SM
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak_tf = 125.0 * 2 * 148 * sm_mhz * 1e6 / 1e12
    traffic_step = load_traffic(args.mode)
    # composite floor: every sub-block is bound by the slowest of HBM, tensor pipe (x products per MAC) and FP32 pipe
    # (depthwise); the sum of those per-layer maxima is the time a perfect kernel of this design would need
    prod = 3 if args.mode == "f16x3" else 1
    if args.mode == "fp32":
        comp_s = sum(max(B * by / (peaks["hbm_gbs"] * 1e9), B * (dwf + pwf) / (fp32_peak_tf * 1e12)) for by, dwf, pwf in counts["layers"])
    else:
        comp_s = sum(max(B * by / (peaks["hbm_gbs"] * 1e9), prod * B * pwf / (peaks["bf16_tflops"] * 1e12),
                         B * dwf / (fp32_peak_tf * 1e12)) for by, dwf, pwf in counts["layers"])
    roofline = {
        "kernel": "segment_kernel / subblock_kernel (fused depthwise + tcgen05 1x1 conv + BN + ReLU; one launch per "
                  "run of same-width sub-blocks and sub-batch)" if args.mode != "fp32"
                  else "dw_conv_kernel + pw_gemm_kernel (CUDA-core path)",
        "bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": ach_gbs / peaks["hbm_gbs"],
        "traffic": (traffic_step / enc_launches) if (traffic_step and enc_launches) else None,
        "peak_source": peaks["source"] + " (burst copy)",
        "what": "all encoder launches of a step taken together: algorithmic bytes of the 78 sub-blocks / encoder-stage time "
                "(CUDA events inside the timed region)",
        "launches_per_step": enc_launches, "sub_blocks_per_step": n_sub,
        "avg_launch_ms": enc_ms_avg / max(1, enc_launches),
        "algorithmic_bytes_per_launch": enc_bytes / max(1, enc_launches),
        "algorithmic_bytes_per_sub_block": enc_bytes / n_sub,
        "tensor": {"achieved": ach_tf, "unit": "TFLOP/s (algorithmic 1x1-conv flops)",
                   "peak": peaks["bf16_tflops"], "frac": ach_tf / peaks["bf16_tflops"],
                   "products_per_mac": 3 if args.mode == "f16x3" else 1,
                   "frac_of_issued_mma": (3 if args.mode == "f16x3" else 1) * ach_tf / peaks["bf16_tflops"],
                   "note": "peak = measured sustained bf16 cuBLAS; f16x3 issues 3 fp16 MMAs per algorithmic MAC"},
        "composite": {"floor_ms": comp_s * 1e3, "frac": comp_s * 1e3 / enc_ms_avg,
                      "note": "sum over the sub-blocks of max(HBM, tensor x products per MAC, FP32 depthwise) floors at the measured "
                              "peaks / encoder-stage time: the fraction of the roofline that actually applies to each layer "
                              "(profiles/r1_layer_table.md)"},
        "fp32_pipe": {"achieved": ach_dw_tf, "unit": "TFLOP/s (depthwise FMAs on the CUDA cores)", "peak": fp32_peak_tf,
                      "frac": ach_dw_tf / fp32_peak_tf,
                      "note": "the co-limiter SURVEY 8(d) names: the depthwise stage runs as packed FFMA2; peak = measured "
                              "125 FMA lanes/clk/SM x 148 SMs x SM clock under load"},
    }
    cpu = None
    if not args.no_cpu_baseline:
        threads = pick_threads(md)
        B_cpu, iters = 16, 2
        v, med = cpu_port_time(md, B_cpu, iters, threads)
        cpu = {"value": v, "unit": "audio-s/s", "cores": threads, "host_cpus": os.cpu_count(), "kind": "port",
               "sample": f"oracle (torch CPU fp32 port of the reference path), {B_cpu} x 5 s clips, median of {iters} passes "
                         f"after 1 warm-up ({med:.2f} s per pass)"}
    line = {
        "metric": "audio-seconds/sec (RTF^-1) QuartzNet15x5 16kHz", "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "p50_ms": statistics.median(step_ms),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "f16x3": "f16x3-split (fp32-grade), fp32 accumulate", "f16x1": "f16, fp32 accumulate"}[args.mode],
        "data": "synthetic",
        "config": {"workload": f"QuartzNet15x5 greedy CTC, batch {B} x 5 s synthetic 16 kHz clips per GPU (BASELINE configs[2])",
                   "global_batch": world * B, "clip_seconds": CLIP_S, "gemm_mode": args.mode, "weights": weights_note,
                   "l2": "per-step working set (82 MB waveforms + 131 MB activations per layer) exceeds the 126 MB L2",
                   "parallelism": f"dp{world}",
                   "e2e_route": ("single C-ABI call vasr_transcribe_host, pinned host buffers" if world == 1 else
                                 ("rank-0 H2D + NCCL scatter + NCCL gather" if args.e2e_scatter else
                                  "per-rank pinned host shards, NCCL gather of ids to rank 0"))},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stage_ms": {"encoder": enc_ms_avg, "step": ms_per_step},
        "latency_b1": lat,
        "legs_agree": legs_agree,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
