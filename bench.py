#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native VietASR CTC hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c2|c4|c5|c5lm] [--mode f16x3|f16x1|fp32]
                    [--impl reference]

A "step" is one pass of the whole hot path (log-mel front end -> QuartzNet encoder -> CTC decoder -> greedy argmax +
CTC collapse, or prefix beam search) over one batch of synthetic 16 kHz clips.  Workloads (BASELINE.json `configs`):

  c3 (default)  configs[2]  QuartzNet15x5, 256 x 5 s per GPU, greedy            weak scaling - the headline line
  c2            configs[1]  QuartzNet12x1 (vi), 32 x 10 s per GPU, greedy       weak scaling
  c4            configs[3]  QuartzNet15x5, 1024 x 5 s over all N GPUs, greedy   STRONG scaling (128 per GPU at N = 8)
  c5            configs[4]  QuartzNet15x5, 128 x 10 s per GPU, beam = 128, no LM (only Vietnamese LMs ship)
  c5lm          configs[4]  QuartzNet12x1 (vi), 128 x 10 s per GPU, beam = 128 + 3-gram KenLM, alpha 0.5, beta 1.5
                            (the reference's default decode, infer.py:184-191)

One JSON line is printed by rank 0 (see the contract in the task statement):
  value      audio-seconds/s, whole job, inputs already resident in HBM, CUDA-event time, max over ranks
  e2e        same metric from HOST buffers: greedy = one C-ABI call per rank (vasr_transcribe_host at N = 1;
             vasr_transcribe_host_to_device + NCCL gather of the ids to rank 0 + D2H there at N > 1; the H2D of each
             rank's pinned shard is pipelined under the compute inside the call); beam = H2D, module calls, D2H
  roofline   dominant kernel family = the fused sub-block kernels of the encoder, timed live with CUDA events around
             the encoder stage inside the timed region
  parity     the CPU port of the reference run on the first clips OF THIS GPU BATCH: ids_equal / logp_rel_l2
  cpu_baseline  the same CPU pass, timed (batched), plus the B = 1 loop infer.py actually runs
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

SR = 16000
CONFIGS = {
    "c3": dict(model="quartznet15x5", weights="en15x5", batch=256, clip_s=5.0, decode="greedy", scaling="weak",
               what="BASELINE configs[2]"),
    "c2": dict(model="quartznet12x1_vi", weights="vi12x1", batch=32, clip_s=10.0, decode="greedy", scaling="weak",
               what="BASELINE configs[1]"),
    "c4": dict(model="quartznet15x5", weights="en15x5", global_batch=1024, clip_s=5.0, decode="greedy", scaling="strong",
               what="BASELINE configs[3] (clip length not named there: 5 s like configs[2])"),
    "c5": dict(model="quartznet15x5", weights="en15x5", batch=128, clip_s=10.0, decode="beam", beam=128, lm=None, scaling="weak",
               what="BASELINE configs[4] without LM (no English LM ships)"),
    "c5lm": dict(model="quartznet12x1_vi", weights="vi12x1", batch=128, clip_s=10.0, decode="beam", beam=128,
                 lm="3-gram-lm.binary", alpha=0.5, beta=1.5, scaling="weak",
                 what="BASELINE configs[4] on the Vietnamese model with its 3-gram KenLM (infer.py:184-191)"),
}
MODEL_NAME = {"quartznet15x5": "QuartzNet15x5", "quartznet12x1_vi": "QuartzNet12x1"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "bf16_tflops_burst": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "source": "fallback"}


def load_traffic(mode):
    """dram__bytes_read.sum + dram__bytes_write.sum of the encoder kernels of ONE step from the committed ncu
    capture (profiles/r2_traffic.json, f16x3, same workload); None for other modes."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
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


def synth_batch(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    wave = (0.1 * torch.randn(B, L, generator=g)).clamp_(-1.0, 1.0)
    length = torch.full((B,), L, dtype=torch.int64)
    return wave, length


def weight_paths(cfg):
    wdir = os.path.join(ROOT, "weights", cfg["weights"])
    e, d = os.path.join(wdir, "JasperEncoder.pt"), os.path.join(wdir, "JasperDecoderForCTC.pt")
    return (e, d) if os.path.exists(e) and os.path.exists(d) else (None, None)


def load_state_dicts(cfg, md):
    from oracle import quartznet_oracle as O
    e, d = weight_paths(cfg)
    if e:
        return torch.load(e, map_location="cpu"), torch.load(d, map_location="cpu"), f"shipped {MODEL_NAME[cfg['model']]} checkpoint"
    enc_sd, dec_sd = O.random_state_dicts(md["JasperEncoder"]["jasper"], 64, len(md["labels"]), seed=1)
    return enc_sd, dec_sd, f"random-init weights of the {MODEL_NAME[cfg['model']]} architecture"


def pick_threads(md, L):
    """The reference path is many small convolutions: torch's intra-op pool scales poorly past a few
    dozen threads.  Probe a short pass at several pool sizes and keep the fastest (all cores are
    available to it; `cores` in the JSON is the pool size actually used)."""
    from oracle import quartznet_oracle as O
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu})
    jasper = md["JasperEncoder"]["jasper"]
    enc_sd, dec_sd = O.random_state_dicts(jasper, 64, len(md["labels"]), seed=1)
    wave, length = synth_batch(4, L, 99)
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


def cpu_port_pass(cfg, md, enc_sd, dec_sd, wave, length, threads, iters, b1_clips):
    """Oracle (port of the reference's torch-CPU path) on the host cores over `wave`: results of the pass (the parity
    check of the GPU batch uses them), batched audio-seconds/s, and the B = 1 loop infer.py:167-171 actually runs."""
    from oracle import quartznet_oracle as O
    torch.set_num_threads(threads)
    jasper = md["JasperEncoder"]["jasper"]
    blank = len(md["labels"])
    clip_s = wave.shape[1] / SR

    def one(w, ln):
        r = O.full_path(enc_sd, dec_sd, jasper, w, ln)
        r["collapsed"] = O.ctc_collapse(r["ids"].numpy(), blank)
        return r

    res = one(wave, length)                                        # warm-up pass; its results are the parity reference
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter(); one(wave, length); ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    t0 = time.perf_counter()
    for i in range(b1_clips):
        one(wave[i:i + 1], length[i:i + 1])
    b1 = (time.perf_counter() - t0) / max(1, b1_clips)
    return res, wave.shape[0] * clip_s / med, med, clip_s / b1, b1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    sys.path.insert(0, os.path.join(ROOT, "viet-asr_b200"))
    import configs as cfgs      # plain module import: the reference arm never loads the CUDA library
    md = cfgs.MODELS[cfg["model"]]()
    from oracle import quartznet_oracle as O
    L = int(cfg["clip_s"] * SR)
    threads = pick_threads(md, L)
    B_cpu = max(1, int(80.0 / cfg["clip_s"]))                     # bounded sample: 80 audio-seconds per step
    steps, warm = max(1, args.steps), max(0, args.warmup)
    torch.set_num_threads(threads)
    jasper = md["JasperEncoder"]["jasper"]
    enc_sd, dec_sd, _ = load_state_dicts(cfg, md)
    wave, length = synth_batch(B_cpu, L, 1234)
    blank = len(md["labels"])
    beam = None
    if cfg["decode"] == "beam":
        # the reference decodes one utterance per call on the CPU (beam_search_decoder.py:95-102); the restatement is
        # pure Python, so the reference arm runs it on ONE clip per step and scales the time to the sample
        from oracle import beam_oracle as BO
        beam = BO
    budget_s = 150.0
    t_start = time.perf_counter()
    ts = []
    for i in range(warm + steps):
        t0 = time.perf_counter()
        r = O.full_path(enc_sd, dec_sd, jasper, wave, length)
        if beam is None:
            O.ctc_collapse(r["ids"].numpy(), blank)
            dt = time.perf_counter() - t0
        else:
            t1 = time.perf_counter()
            beam.beam_search_no_lm(r["logp"][0].numpy(), md["labels"], cfg["beam"])
            dt = (t1 - t0) + (time.perf_counter() - t1) * B_cpu
        if i >= warm:
            ts.append(dt)
        if time.perf_counter() - t_start > budget_s and len(ts) >= 1:
            break
    med = statistics.median(ts)
    v = B_cpu * cfg["clip_s"] / med
    sample = (f"{B_cpu} x {cfg['clip_s']:.0f} s synthetic clips per step ({len(ts)} timed steps), torch CPU fp32"
              + ("; beam search (no LM, Python restatement of pyctcdecode) timed on one clip and scaled to the sample" if beam else ""))
    print(json.dumps({
        "impl": "reference", "metric": f"audio-seconds/sec (RTF^-1) {MODEL_NAME[cfg['model']]} 16kHz", "value": v,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": len(ts), "warmup": warm, "ms_per_step": med * 1e3,
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {MODEL_NAME[cfg['model']]} {cfg['decode']} CTC, {cfg['clip_s']:.0f} s synthetic 16 kHz clips "
                               f"({cfg['what']}; CPU arm: bounded sample of {B_cpu} clips per step)"},
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads, "host_cpus": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="f16x3", choices=["fp32", "f16x3", "f16x1"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch of the config")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU pass (no cpu_baseline / parity fields)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import viet_asr_b200 as V
    from viet_asr_b200 import dist as D

    cfg = CONFIGS[args.config]
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
    CLIP_S = cfg["clip_s"]
    L = int(CLIP_S * SR)
    if args.batch > 0:
        B = args.batch
    elif "global_batch" in cfg:
        B = cfg["global_batch"] // world                         # strong scaling: the job's batch is fixed
    else:
        B = cfg["batch"]
    GB = world * B
    beam = cfg["decode"] == "beam"

    md = V.configs.MODELS[cfg["model"]]()
    jasper = md["JasperEncoder"]["jasper"]
    labels = md["labels"]
    V.NeuralModuleFactory(placement=V.DeviceType.GPU)
    lm_path = None
    if beam and cfg.get("lm"):
        lm_path = os.path.join(ROOT, "weights", "lm", cfg["lm"])
        if not os.path.exists(lm_path):
            raise SystemExit(f"{lm_path} is missing (oracle/make_golden.py --lm-only copies it where /root/reference exists)")
    eng = V.VietASR(model_definition=md, gemm_mode=args.mode, lm_path=lm_path, beam_width=cfg.get("beam", 20),
                    lm_alpha=cfg.get("alpha", 0.5), lm_beta=cfg.get("beta", 1.5))
    enc_sd, dec_sd, weights_note = load_state_dicts(cfg, md)
    eng.load_state_dicts(enc_sd, dec_sd)
    blank = len(labels)

    wave_h, len_h = synth_batch(B, L, 1234 + rank)
    wave_d, len_d = wave_h.to(dev), len_h.to(dev)
    T_f = eng.preprocessor.num_frames(L)
    T_e = eng.out_frames(L)
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
    def step_device(marks=None, want_log_probs=False):
        feat, seq = eng.preprocessor.forward_channels_last(wave_d, len_d)
        if marks: marks[0].record()
        enc, enc_len = eng.encoder.forward_channels_last(feat, seq)
        if marks: marks[1].record()
        logp, ids = eng.decoder.forward_channels_last(enc, beam or want_log_probs)
        if marks: marks[2].record()
        if beam:
            out_ids, out_len, _ = V.ctc_beam_search(logp, labels, cfg["beam"], lm=eng.beam.lm, alpha=cfg.get("alpha", 0.5),
                                                    beta=cfg.get("beta", 1.5))
        else:
            out_ids, out_len = V.ctc_collapse(ids, blank)
        return out_ids, out_len, ids, logp

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
    step_ev = [tuple(ev() for _ in range(5)) for _ in range(K)]
    barrier()
    for i in range(K):
        s, e0, e1, e2, e = step_ev[i]
        s.record()
        step_device((e0, e1, e2))
        e.record()
    barrier()
    launches = V._lib.launch_count() - launches0
    total_ms = max_over_ranks(step_ev[0][0].elapsed_time(step_ev[-1][4]))
    step_ms = [a.elapsed_time(d) for a, _, _, _, d in step_ev]
    ms_per_step = total_ms / K
    value = GB * CLIP_S / (ms_per_step * 1e-3)
    fe_ms_avg = max_over_ranks(sum(a.elapsed_time(b) for a, b, _, _, _ in step_ev) / K)
    enc_ms_avg = max_over_ranks(sum(b.elapsed_time(c) for _, b, c, _, _ in step_ev) / K)
    dec_ms_avg = max_over_ranks(sum(c.elapsed_time(d) for _, _, c, d, _ in step_ev) / K)
    search_ms_avg = max_over_ranks(sum(d.elapsed_time(e) for _, _, _, d, e in step_ev) / K)

    # ------------------------------------------------------------ leg 2: end to end, host buffers
    wave_p, len_p = wave_h.pin_memory(), len_h.pin_memory()
    if rank == 0:
        oid_p = torch.empty((GB, T_e), dtype=torch.int32).pin_memory()
        oln_p = torch.empty((GB,), dtype=torch.int32).pin_memory()
    w_dev = torch.empty((B, L), dtype=torch.float32, device=dev) if beam else None
    l_dev = torch.empty((B,), dtype=torch.int64, device=dev) if beam else None
    oid_d = torch.empty((B, T_e), dtype=torch.int32, device=dev)
    oln_d = torch.empty((B,), dtype=torch.int32, device=dev)

    def step_e2e():
        if beam:
            # host shard -> H2D -> front end .. decoder -> beam search -> ids (module calls of the public API)
            w_dev.copy_(wave_p, non_blocking=True); l_dev.copy_(len_p, non_blocking=True)
            feat, seq = eng.preprocessor.forward_channels_last(w_dev, l_dev)
            enc, _ = eng.encoder.forward_channels_last(feat, seq)
            logp, _ = eng.decoder.forward_channels_last(enc, True)
            ids_d, n_d, _ = V.ctc_beam_search(logp, labels, cfg["beam"], lm=eng.beam.lm, alpha=cfg.get("alpha", 0.5),
                                              beta=cfg.get("beta", 1.5))
        elif world == 1:
            eng.transcribe_host_ids(wave_p, len_p, oid_p, oln_p)          # ONE C-ABI call: H2D pipelined under compute, D2H, sync
            return
        else:
            # every rank: its own pinned shard through vasr_transcribe_host_to_device (copy pipelined under compute, the
            # ids stay on the device), then the NCCL gather to rank 0 and the D2H there
            ids_d, n_d = eng.transcribe_host_to_device(wave_p, len_p, oid_d, oln_d)
        if world > 1:
            gi, gn = D.gather_results(ids_d, n_d, GB)
        else:
            gi, gn = ids_d, n_d
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
    e2e_value = GB * CLIP_S / (e2e_ms * 1e-3)
    clocks = sampler.stop() if sampler else None
    if not beam:
        eng.check_range() if world > 1 else None

    # ------------------------------------------------------------ result identity across legs and ranks
    o_ids, o_len, f_ids, f_logp = step_device(want_log_probs=True)
    if world > 1:
        gi, gn = D.gather_results(o_ids, o_len, GB)                       # device-leg results of every rank, on rank 0
    else:
        gi, gn = o_ids, o_len
    legs_agree = ranks_agree = None
    if rank == 0:
        # the transcripts the e2e leg delivered to rank 0's host buffers == what every rank's device leg computes
        legs_agree = bool(torch.equal(gi.cpu(), oid_p) and torch.equal(gn.cpu(), oln_p))
        if world > 1:
            # 1-GPU identity: rank 0 recomputes the LAST rank's batch on its own GPU and compares with what that rank sent
            w_last, l_last = synth_batch(B, L, 1234 + world - 1)
            r_last = eng.forward_device(w_last.to(dev), l_last.to(dev)) if not beam else None
            if r_last is not None:
                n = B
                ranks_agree = bool(torch.equal(r_last["out_ids"].cpu(), oid_p[GB - n:]) and torch.equal(r_last["out_len"].cpu(), oln_p[GB - n:]))
        if not legs_agree:
            print("WARNING: device-resident and host-buffer legs produced different transcripts", file=sys.stderr)

    # ------------------------------------------------------------ single-utterance latency (the web-app case)
    lat = None
    if rank == 0 and not beam:
        w1, l1 = synth_batch(1, L, 777)
        w1p, l1p = w1.pin_memory(), l1.pin_memory()
        o1 = torch.empty((1, T_e), dtype=torch.int32).pin_memory(); n1 = torch.empty((1,), dtype=torch.int32).pin_memory()
        for _ in range(5):
            eng.transcribe_host_ids(w1p, l1p, o1, n1)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); eng.transcribe_host_ids(w1p, l1p, o1, n1); ts.append((time.perf_counter() - t0) * 1e3)
        lat = {"batch": 1, "clip_seconds": CLIP_S, "p50_ms": statistics.median(ts), "p90_ms": sorted(ts)[26],
               "what": "vasr_transcribe_host wall time (H2D + whole path + D2H + sync), greedy"}
        try:
            gr = eng.capture_graph(1, L)
            for _ in range(5):
                gr(w1p, l1p)
            ts = []
            for _ in range(30):
                t0 = time.perf_counter(); gr(w1p, l1p); ts.append((time.perf_counter() - t0) * 1e3)
            same = bool(torch.equal(gr(w1p, l1p)[0][:, : o1.shape[1]], o1))
            lat["graph"] = {"p50_ms": statistics.median(ts), "p90_ms": sorted(ts)[26], "equals_call_route": same,
                            "what": "the same path captured once as a CUDA graph (VietASR.capture_graph): H2D + graph launch + D2H + sync"}
        except Exception as e:                                               # the graph route is an extra, never the measured path
            lat["graph"] = {"error": str(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------ CPU port on the first clips OF THIS BATCH: parity + baseline
    cpu = parity = None
    if not args.no_cpu_baseline:
        threads = pick_threads(md, L)
        n_par = max(2, int(80.0 / CLIP_S)) if not beam else max(2, int(40.0 / CLIP_S))
        n_par = min(n_par, B)
        ref, v, med, v_b1, t_b1 = cpu_port_pass(cfg, md, enc_sd, dec_sd, wave_h[:n_par], len_h[:n_par], threads, 2, min(4, n_par))
        ids_equal = bool(torch.equal(f_ids[:n_par].cpu(), ref["ids"]))
        rel = ((f_logp[:n_par].cpu() - ref["logp"]).norm() / ref["logp"].norm()).item()
        t2 = ref["logp"].topk(2, -1).values
        near_tie = (t2[..., 0] - t2[..., 1]) <= 1e-3               # the reference's own two best classes within the tolerance
        differ = f_ids[:n_par].cpu() != ref["ids"]
        parity = {"clips": n_par, "ids_equal": ids_equal, "logp_rel_l2": rel, "frames": int(differ.numel()),
                  "frames_differing": int(differ.sum()), "frames_differing_outside_near_ties": int((differ & ~near_tie).sum()),
                  "near_tie_frames": int(near_tie.sum()),
                  "what": f"CPU port of the reference (oracle/quartznet_oracle.py) on the first {n_par} clips of rank 0's GPU batch vs rows "
                          f"[0, {n_par}) of the device leg at batch {B}: greedy ids bit-exact, log-probs rel-L2 (bound 1e-3)"}
        if beam:
            from oracle import beam_oracle as BO
            if eng.beam.lm is None:
                want, _ = BO.beam_search_no_lm(ref["logp"][0].numpy(), labels, cfg["beam"])
            else:
                from oracle.kenlm_oracle import KenlmBinary
                want, _ = BO.beam_search_lm(ref["logp"][0].numpy(), labels, cfg["beam"], KenlmBinary(lm_path), alpha=cfg["alpha"], beta=cfg["beta"])
            got = " ".join(V.ids_to_text(o_ids[:1], o_len[:1], labels)[0].split())
            parity["beam_text_equal_clip0"] = bool(got == want)
            parity["beam_note"] = "beam search vs oracle/beam_oracle.py (restatement of pyctcdecode: parity unpinned against the package)"
        else:
            parity["collapsed_equal"] = bool([r[:k].tolist() for r, k in zip(o_ids[:n_par].cpu().numpy(), o_len[:n_par].cpu().numpy())] == ref["collapsed"])
        cpu = {"value": v, "unit": "audio-s/s", "cores": threads, "host_cpus": os.cpu_count(), "kind": "port",
               "sample": f"oracle (torch CPU fp32 port of the reference path, greedy), the first {n_par} x {CLIP_S:.0f} s clips of the GPU batch, "
                         f"median of 2 passes after 1 warm-up ({med:.2f} s per pass)",
               "b1_loop": {"value": v_b1, "unit": "audio-s/s", "s_per_clip": t_b1,
                           "what": "one clip per call like infer.py:167-171 / app.py (how the reference actually runs), mean of "
                                   f"{min(4, n_par)} clips"}}

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
    traffic_step = load_traffic(args.mode) if args.config == "c3" else None
    prod = 3 if args.mode == "f16x3" else 1
    # the bench runs at 1.8-1.97 GHz (clocks field), where cuBLAS reaches its BURST figure; the sustained figure of
    # MEASURED_PEAKS.json was taken at 1.33 GHz under a 1 kW load this kernel does not draw
    tc_peak = peaks["bf16_tflops_burst"]
    if args.mode == "fp32":
        comp_s = sum(max(B * by / (peaks["hbm_gbs"] * 1e9), B * (dwf + pwf) / (fp32_peak_tf * 1e12)) for by, dwf, pwf in counts["layers"])
    else:
        comp_s = sum(max(B * by / (peaks["hbm_gbs"] * 1e9), prod * B * pwf / (tc_peak * 1e12),
                         B * dwf / (fp32_peak_tf * 1e12)) for by, dwf, pwf in counts["layers"])
    roofline = {
        "kernel": "segment_pair_kernel / segment_kernel / subblock_kernel (fused depthwise + tcgen05 1x1 conv + BN + ReLU; one launch per "
                  "run of same-width sub-blocks and sub-batch)" if args.mode != "fp32"
                  else "dw_conv_kernel + pw_gemm_kernel (CUDA-core path)",
        "bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": ach_gbs / peaks["hbm_gbs"],
        "traffic": (traffic_step / enc_launches) if (traffic_step and enc_launches) else None,
        "peak_source": peaks["source"] + " (burst copy)",
        "what": f"all encoder launches of a step taken together: algorithmic bytes of the {n_sub} sub-blocks / encoder-stage time "
                "(CUDA events inside the timed region)",
        "launches_per_step": enc_launches, "sub_blocks_per_step": n_sub,
        "avg_launch_ms": enc_ms_avg / max(1, enc_launches),
        "algorithmic_bytes_per_launch": enc_bytes / max(1, enc_launches),
        "algorithmic_bytes_per_sub_block": enc_bytes / n_sub,
        "tensor": {"achieved": ach_tf, "unit": "TFLOP/s (algorithmic 1x1-conv flops)",
                   "peak": tc_peak, "frac": ach_tf / tc_peak, "peak_sustained": peaks["bf16_tflops"],
                   "products_per_mac": prod, "frac_of_issued_mma": prod * ach_tf / tc_peak,
                   "note": "peak = measured burst bf16 cuBLAS (the bench runs near the maximum SM clock, see `clocks`); f16x3 issues 3 fp16 "
                           "MMAs per algorithmic MAC; the ncu tensor-pipe counter of the same kernels is in profiles/"},
        "composite": {"floor_ms": comp_s * 1e3, "frac": comp_s * 1e3 / enc_ms_avg,
                      "note": "sum over the sub-blocks of max(HBM, tensor x products per MAC, FP32 depthwise) floors at the measured "
                              "peaks / encoder-stage time: the fraction of the roofline that actually applies to each layer"},
        "fp32_pipe": {"achieved": ach_dw_tf, "unit": "TFLOP/s (depthwise FMAs on the CUDA cores)", "peak": fp32_peak_tf,
                      "frac": ach_dw_tf / fp32_peak_tf,
                      "note": "the co-limiter SURVEY 8(d) names: the depthwise stage runs as packed FFMA2; peak = measured "
                              "125 FMA lanes/clk/SM x 148 SMs x SM clock under load"},
    }
    stage_ms = {"frontend": fe_ms_avg, "encoder": enc_ms_avg, "decoder": dec_ms_avg,
                ("beam_search" if beam else "collapse"): search_ms_avg, "step": ms_per_step}
    decode_info = None
    if beam:
        decode_info = {"kernel": "beam_kernel<LM>" if eng.beam.lm is not None else "beam_kernel<no LM>", "ms": search_ms_avg,
                       "share_of_step": search_ms_avg / ms_per_step, "beam_width": cfg["beam"], "frames": T_e,
                       "bound": "latency: one CTA per utterance (two per SM), T_e strictly dependent frames of 6-9 block barriers each "
                                "(candidate expansion, hash-table merge in shared memory, prune, radix select + rank by counting)",
                       "us_per_frame": search_ms_avg * 1e3 / T_e}
    line = {
        "metric": f"audio-seconds/sec (RTF^-1) {MODEL_NAME[cfg['model']]} 16kHz", "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "p50_ms": statistics.median(step_ms),
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": {"fp32": "f32", "f16x3": "f16x3-split (fp32-grade), fp32 accumulate", "f16x1": "f16, fp32 accumulate"}[args.mode],
        "data": "synthetic",
        "config": {"workload": f"{args.config}: {MODEL_NAME[cfg['model']]} {cfg['decode']} CTC"
                               + (f" (beam {cfg['beam']}{', ' + cfg['lm'] if cfg.get('lm') else ', no LM'})" if beam else "")
                               + f", batch {B} x {CLIP_S:.0f} s synthetic 16 kHz clips per GPU ({cfg['what']})",
                   "global_batch": GB, "clip_seconds": CLIP_S, "gemm_mode": args.mode, "weights": weights_note,
                   "l2": f"per-step working set ({B * L * 4 / 1e6:.0f} MB waveforms + {B * T_e * 512 * 4 / 1e6:.0f} MB activations per "
                         "layer) exceeds the 126 MB L2" if B * L * 4 + B * T_e * 2048 > 126e6 else
                         "working set fits the 126 MB L2: a 268 MB buffer is NOT flushed between steps (latency-style config)",
                   "parallelism": f"dp{world}",
                   "e2e_route": ("H2D + module calls + beam search + D2H" if beam else
                                 "single C-ABI call vasr_transcribe_host, pinned host buffers" if world == 1 else
                                 "per-rank pinned host shards through vasr_transcribe_host_to_device (pipelined H2D), NCCL gather of ids to rank 0, D2H")},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "stage_ms": stage_ms,
        "decode_stage": decode_info,
        "latency_b1": lat,
        "legs_agree": legs_agree,
        "ranks_agree": ranks_agree,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
