#!/usr/bin/env python
"""Benchmark of the audio -> LSM-feature hot path (BASELINE.json metric: utterances/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one pass of the whole path (gammatone filterbank + hysteresis encoder -> reservoir ->
feature readout) over one batch of synthetic 1 s / 16 kHz utterances: BASELINE.json configs[1]
(12 classes, 2400 utterances, 128-channel gammatone, N=1000 reservoir, feature set `original`).
One JSON line on stdout (rank 0).  Under torchrun each rank runs its own batch (weak scaling) and the
raw feature rows are all-gathered over NCCL inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CLASSES, PER_CLASS = 12, 200          # configs[1]: 2400 utterances
N_FILTERS, FILTERBANK, FEATURE_SET, MULTIPLIER = 128, "gammatone", "original", 0.6
N_NEURONS, T_STEPS, L = 1000, 400, 16000
# algorithmic bytes per utterance: K1 = 64 000 B PCM in + 51 200 B spikes out; the fused
# audio -> features kernel adds the 16 000 B feature row (2000 x fp64)
K1_BYTES_PER_UTT = 64000 + 51200
FUSED_BYTES_PER_UTT = 64000 + 51200 + 16000
# fp64-pipe operations per utterance in K1: 128 ch x 15920 samples x 35 (28 biquad + 3 divide + 1 square + 3 window adds)
K1_FP64_OPS_PER_UTT = 128 * 15920 * 35
# the speculative arrangement of the same cascade (default mode): 12 FMAs for the four sections + 1 for the energy sum
K1_FP64_OPS_PER_UTT_SPEC = 128 * 15920 * 13
METRIC = "utterances/sec audio->LSM features"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # under load = the upper half of the samples (idle tail/head excluded)
            sm_sorted = sorted(sm)
            out["sm_mhz"] = statistics.median(sm_sorted[len(sm_sorted) // 2:])
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def make_inputs(rank: int):
    from lsm_speech_classifier_b200 import synth
    t0 = time.time()
    # each rank gets different utterances (utt ids offset by rank): weak scaling, fixed work per GPU
    pcm, labels = synth.synth_dataset(N_CLASSES, PER_CLASS, start_utt=rank * PER_CLASS, workers=os.cpu_count() or 1)
    log(f"[bench] rank {rank}: synthesised {len(pcm)} utterances in {time.time() - t0:.1f}s")
    return pcm, labels


def oracle_pipeline(pcm, fe_tables, res, keys_mask, nthreads):
    from oracle import coracle
    table, nwin, hop, nbins, zi0, zf = fe_tables
    spikes = coracle.gammatone_encode(pcm, table, nwin, hop, nbins, zi0, zf, [0.70, 0.80, 0.90, 0.95], 0.1, nthreads=nthreads)
    feats, _ = coracle.reservoir_run(res, spikes, keys_mask, True, False, nthreads=nthreads)
    return feats


def host_tables():
    from lsm_speech_classifier_b200 import filterbank as fb
    nwin, hop, ncols = fb.gtgram_strides(16000, 0.025, 0.01, L)
    zi0, zf = fb.zoom_table(ncols, 100)
    return fb.gammatone_coefs(16000, N_FILTERS, 50), nwin, hop, 100, zi0, zf


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path, timed on the host cores.  The reference
    (pure Python delegating to gammatone/librosa/snnpy, none installable here) cannot run, so this is
    the oracle port (plain C, one pthread per core over utterances) on a bounded sample per step."""
    if rank != 0:
        return
    from oracle import coracle
    from oracle import pyref
    from lsm_speech_classifier_b200 import synth
    from lsm_speech_classifier_b200.reservoir import SimulationParams, build_reservoir
    from lsm_speech_classifier_b200._lib import feature_mask
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS
    cores = coracle.num_threads()
    # the whole 2400-utterance step when the host gets through it in a few seconds (>= 16 threads), else a bounded sample
    per_class = PER_CLASS if cores >= 16 else max(1, (cores * 8 + N_CLASSES - 1) // N_CLASSES)
    pcm, _ = synth.synth_dataset(N_CLASSES, per_class, workers=cores)
    tables = host_tables()
    spikes = coracle.gammatone_encode(pcm[:64], *tables, [0.70, 0.80, 0.90, 0.95], 0.1)
    wc = pyref.w_critico(200, 2.0, 2, list(spikes))
    res = build_reservoir(SimulationParams(mean_weight=wc * MULTIPLIER, input_spike_times=spikes[0]))
    mask = feature_mask(FEATURE_SETS[FEATURE_SET])
    for _ in range(args.warmup):
        oracle_pipeline(pcm, tables, res, mask, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_pipeline(pcm, tables, res, mask, 0)
    dt = time.perf_counter() - t0
    val = len(pcm) * args.steps / dt
    sample = (f"the full {len(pcm)}-utterance step, all host threads" if len(pcm) == N_CLASSES * PER_CLASS else
              f"{len(pcm)} utterances per step (bounded sample of the {N_CLASSES * PER_CLASS}-utterance workload), all host threads")
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "utterances/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": val, "unit": "utterances/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "neuron_steps_per_s": val * N_NEURONS * T_STEPS,
    })


def workload_config():
    return {"workload": "configs[1]: 12-class, 2400 synthetic 1 s/16 kHz utterances per step per GPU, 128-ch gammatone, "
                        "LSM N=1000 k=200 T=400, feature set original (2000 features), multiplier 0.6",
            "utterances_per_step_per_gpu": N_CLASSES * PER_CLASS, "n_filters": N_FILTERS, "filterbank": FILTERBANK,
            "feature_set": FEATURE_SET, "multiplier": MULTIPLIER, "n_neurons": N_NEURONS,
            "l2_policy": "inputs larger than L2 (153.6 MB PCM per step > 126 MB L2)",
            "parallelism": "utterance-sharded, one process per GPU; feature rows all-gathered inside the timed region, "
                           + ("over NCCL" if os.environ.get("LSM_BENCH_NCCL_GATHER")
                              else "by the kernel's readout epilogue (NVLink stores into every rank's matrix; no collective kernel)")}


# mel front end (BASELINE.json configs[2]): fp64 lane-operations per utterance in K1m - per frame 2048 window products, 5120
# radix-2 butterflies x 10, 1025 untangle + magnitude bins x 20; 101 frames; + ~45 per dB value of the 101 x C plane
MEL_FP64_OPS_PER_UTT = 101 * (2048 + 5120 * 10 + 1025 * 20) + 101 * 128 * 45


def mel_numbers(pcm_np, steps, warmup, cpu_sample=96):
    """The same step through the mel front end (librosa branch of create_dataset.py:43-48), one GPU: the fused
    audio -> features path (spike trains stay on chip), K1m and K2 alone, host buffers end to end, parity with the C oracle on a sample."""
    import torch
    from lsm_speech_classifier_b200 import _lib, filterbank as fb
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    from oracle import coracle
    keys = FEATURE_SETS[FEATURE_SET]
    B = len(pcm_np)
    fe = Frontend(N_FILTERS, "mel")
    d_pcm = torch.from_numpy(pcm_np).cuda()
    head = fe.encode(d_pcm[:500]).cpu().numpy()
    lsm = build_lsm(head, MULTIPLIER, verbose=False)
    path = AudioToFeatures(fe, lsm)
    F = len(keys) * lsm.num_output_neurons
    d_feat = torch.empty((B, F), dtype=torch.float64, device="cuda")
    d_spk = fe.encode(d_pcm)

    def timed(fn, reps):
        for _ in range(max(1, warmup)):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    step_ms = timed(lambda: path.run(d_pcm, keys, out=d_feat, want_spikes=False), steps)
    k1_ms = timed(lambda: fe.encode(d_pcm), max(2, steps // 2))
    k2_ms = timed(lambda: lsm.simulate_batch(d_spk, keys), max(2, steps // 2))
    out_np = np.empty((B, F), dtype=np.float64)
    path.run_host(pcm_np, keys, out=out_np)
    t0 = time.perf_counter()
    n_e2e = max(2, min(steps, 5))
    for _ in range(n_e2e):
        path.run_host(pcm_np, keys, out=out_np)
    e2e = B * n_e2e / (time.perf_counter() - t0)
    same_as_device = bool(np.array_equal(out_np, d_feat.cpu().numpy()))
    sample = np.ascontiguousarray(pcm_np[:: max(1, B // cpu_sample)][:cpu_sample])
    t0 = time.perf_counter()
    spk = coracle.mel_encode(sample, fe.table, fe.window, fe.tw, fe.tw2, fb.pack_mel_basis(fe.table), fe.params.mel_hop, fe.time_bins,
                             fe.zoom_i0, fe.zoom_f, [0.70, 0.80, 0.90, 0.95], 0.1)
    want, _ = coracle.reservoir_run(lsm.reservoir, spk, _lib.feature_mask(keys), True, False)
    cpu = len(sample) / (time.perf_counter() - t0)
    got = path.run_host(sample, keys)
    peak = fe.ctx.fp64_peak_gops()
    gops = MEL_FP64_OPS_PER_UTT * B / (k1_ms / 1e3) / 1e9
    return {"filterbank": "mel", "value": B / (step_ms / 1e3), "unit": "utterances/s", "ms_per_step": step_ms, "fused_one_kernel": bool(path.fused),
            "kernel_ms": {"K1m_mel_encode": k1_ms, "K2_reservoir_features": k2_ms},
            "e2e": {"value": e2e, "unit": "utterances/s", "note": "pageable numpy arrays through lsm_pipeline_run_host (pinned ring filled by host copy threads | H2D | kernels | D2H on two lanes)",
                    "same_rows_as_device_path": same_as_device},
            "roofline": {"kernel": "mel_power_kernel + mel_finish_kernel (K1m)", "bound": "fp64", "achieved": gops, "peak": peak,
                         "unit": "G fp64 lane-ops/s", "frac": gops / peak, "lane_ops_per_utterance": MEL_FP64_OPS_PER_UTT,
                         "note": "warp-per-frame fp64 radix-2 FFT in registers: 8 warps per SM (250 registers per lane), latency / instruction-fetch bound, not pipe bound (DESIGN.md K1m)"},
            "cpu_baseline": {"value": cpu, "unit": "utterances/s", "cores": coracle.num_threads(), "kind": "port",
                             "sample": f"{len(sample)} utterances, oracle C port (mel)", "gpu_matches_cpu_bit_exact": bool(np.array_equal(got, want))}}


_REAL_STDOUT = None


def _claim_stdout():
    """Libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the whole run and keep the real stdout for that line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-times", action="store_true", help="(kept for compatibility; per-kernel times are always reported)")
    ap.add_argument("--filterbank", type=str, default="gammatone", choices=["gammatone", "mel"],
                    help="mel: the same step through the mel front end (one GPU), printed as its own line")
    ap.add_argument("--filter-mode", type=str, default="speculative", choices=["speculative", "exact"],
                    help="gammatone filter evaluation (include/lsm_b200.h): both give the reference-order spike trains")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    # synthesise the inputs first: the generator forks worker processes, which must happen before CUDA is initialised
    pcm_np, _ = make_inputs(rank)
    if args.filterbank == "mel":
        if rank == 0:
            m = mel_numbers(pcm_np, args.steps, max(args.warmup, 3))
            cfg = workload_config()
            cfg["workload"] = cfg["workload"].replace("128-ch gammatone", "128-ch mel (configs[2] front end)")
            cfg["filterbank"] = "mel"
            emit({"metric": METRIC, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True,
                  "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic", "config": cfg, **m})
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from lsm_speech_classifier_b200 import _lib
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures

    keys = FEATURE_SETS[FEATURE_SET]
    B = len(pcm_np)
    h_pcm = torch.from_numpy(pcm_np).pin_memory()
    d_pcm = h_pcm.cuda(non_blocking=True)
    fe = Frontend(N_FILTERS, FILTERBANK)
    fe.set_mode(args.filter_mode)
    ctx = fe.ctx
    # w_critico from the first <=500 spike trains (extract_lsm_features.py:40-44), then the one reservoir
    head = fe.encode(d_pcm[:500]).cpu().numpy()
    lsm = build_lsm(head, MULTIPLIER, verbose=False)
    path = AudioToFeatures(fe, lsm)
    F = len(keys) * lsm.num_output_neurons
    # Two buffer sets and two streams: step i runs on stream i & 1.  The front end has two scratch slots, so two launches are
    # in flight and the drain tail of one batch (whole-utterance granularity) overlaps the start of the next; under torchrun
    # the all-gather of step i (NCCL's stream) also overlaps the kernel of step i+1.
    d_spikes2 = [torch.empty((B, fe.rows, fe.steps), dtype=torch.uint8, device="cuda") for _ in range(2)]
    d_spikes = d_spikes2[0]
    d_feats = [torch.empty((B, F), dtype=torch.float64, device="cuda") for _ in range(2)]
    d_feat = d_feats[0]
    # Feature all-gather (the path's one collective).  Default: fused into the readout epilogue - every rank's kernel stores its
    # feature rows into all ranks' gather matrices itself (lsm_reservoir_set_gather; the peers' matrices are mapped through CUDA
    # IPC), so no collective kernel has to find room beside the persistent kernels.  Checked against one NCCL all-gather after the
    # timed regions.  LSM_BENCH_NCCL_GATHER=1: NCCL's asynchronous all-gather instead; LSM_BENCH_P2P=1: copy-engine peer copies.
    fused_gather = world > 1 and not os.environ.get("LSM_BENCH_NCCL_GATHER") and not os.environ.get("LSM_BENCH_P2P")
    use_p2p = world > 1 and bool(os.environ.get("LSM_BENCH_P2P") or fused_gather)
    pag = None
    if use_p2p:
        from lsm_speech_classifier_b200.distributed import PeerAllGather
        try:
            pag = PeerAllGather(B, F, torch.float64, torch.device("cuda", local_rank), ctx=ctx)
        except RuntimeError as e:        # raised on every rank alike: no peer access here, the collective goes over NCCL
            log(f"[bench] rank {rank}: {e}; feature all-gather over NCCL")
            pag, fused_gather, use_p2p = None, False, False
            os.environ["LSM_BENCH_NCCL_GATHER"] = "1"
    if pag is not None:
        d_alls = pag.bufs
    else:
        d_alls = [torch.empty((world * B, F), dtype=torch.float64, device="cuda") for _ in range(2)] if world > 1 else None
    d_all = d_alls[0] if world > 1 else None
    h_feats = [torch.empty((B, F), dtype=torch.float64).pin_memory() for _ in range(2)]
    h_feat = h_feats[0]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    pending = [None, None]
    step_no = [0]

    def step_device():
        b = step_no[0] & 1
        step_no[0] += 1
        with torch.cuda.stream(streams[b]):
            if pending[b] is not None:
                pending[b].wait()          # the all-gather that last read this buffer pair
            if pag is not None and not fused_gather:
                pag.wait(b)                # this rank's copies out of d_feats[b] two steps ago
            if fused_gather:
                lsm.set_gather(pag.pointers(b), rank * B)      # read when the launch is enqueued
            path.run(d_pcm, keys, spikes=d_spikes2[b], out=d_feats[b])
            if pag is not None and not fused_gather:
                pag.gather_async(b, d_feats[b], streams[b])
            elif world > 1 and not fused_gather:
                pending[b] = dist.all_gather_into_tensor(d_alls[b], d_feats[b], async_op=True)

    def drain():
        for b in (0, 1):
            with torch.cuda.stream(streams[b]):
                if pending[b] is not None:
                    pending[b].wait()
                    pending[b] = None
            torch.cuda.current_stream().wait_stream(streams[b])
            if pag is not None:
                pag.wait(b)

    def fence():
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    fence()

    # ---- value: device-resident inputs, CUDA events, max over ranks.  e0 is recorded on the default stream and both
    #      step streams wait for it; e1 is recorded after the default stream has joined both step streams.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record()
    for st in streams:
        st.wait_event(e0)
    for _ in range(args.steps):
        step_device()
    drain()                 # every kernel and all-gather has finished before the closing event
    e1.record()
    fence()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = ctx.launches - launches0 + (args.steps if (world > 1 and pag is None) else 0)      # + NCCL's kernels, if it gathers
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- the same steps for >= 2 s: the sustained figure beside the K-step burst (clocks settle under the power cap)
    n_sus = max(args.steps, int(2200.0 / max(ms_total / args.steps, 1e-3)) + 1)
    fence()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for st in streams:
        st.wait_event(s0)
    for _ in range(n_sus):
        step_device()
    drain()
    s1.record()
    fence()
    ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    sustained_ms = float(ts.item())

    # ---- per-kernel durations (K1 alone, K2 alone) for the roofline: CUDA events, same stream
    def time_kernel(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    if fused_gather:
        lsm.set_gather([], 0)              # the stand-alone timings below do not gather
    reps = max(3, min(args.steps, 10))
    reruns_value = fe.reruns(reset=True) / max(1, max(args.warmup, 3) + args.steps + 1)   # per step (+1: the w_critico head)
    k1_ms = time_kernel(lambda: fe.encode(d_pcm), reps)
    k2_ms = time_kernel(lambda: lsm.simulate_batch(d_spikes, keys), reps)
    # the whole step as back-to-back launches on ONE stream: the duration of a single launch of the fused kernel
    fused_ms = time_kernel(lambda: path.run(d_pcm, keys, spikes=d_spikes, out=d_feats[0]), reps)
    # the other filter mode beside it: stand-alone K1 and the whole fused step, and the two modes' outputs compared
    other = "exact" if args.filter_mode == "speculative" else "speculative"
    feats_this = d_feats[0].clone()
    spikes_this = d_spikes.clone()
    fe.set_mode(other)
    k1_other_ms = time_kernel(lambda: fe.encode(d_pcm), reps)
    step_other_ms = time_kernel(lambda: path.run(d_pcm, keys, spikes=d_spikes, out=d_feats[0]), reps)
    modes_agree = bool(torch.equal(spikes_this, d_spikes) and torch.equal(feats_this, d_feats[0]))
    fe.set_mode(args.filter_mode)
    del feats_this, spikes_this

    # ---- e2e: host buffers through the public call, H2D + D2H inside the timed region.  Pinned buffers: the fused kernel
    #      reads the PCM and writes the feature rows across PCIe itself; consecutive batches alternate the ctx's two launch
    #      lanes (lsm_pipeline_run_host_async), one sync at the end.  Under torchrun the feature rows stay in device memory for
    #      the all-gather, which is ordered after the kernel on the same lane, and the local rows are copied to the host.
    for _ in range(2):
        path.run_host(h_pcm.numpy(), keys, out=h_feat.numpy())
    fence()
    ext = [torch.cuda.ExternalStream(ctx.lane_stream(k)) for k in range(2)] if world > 1 else None

    def e2e_loop(h_in, feed="zero_copy"):
        ctx.set_host_feed(feed)
        pend = [None, None]

        def one(i):
            b = i & 1
            if world == 1:
                path.run_host_async(h_in, keys, out=h_feats[b], lane=b)    # pinned in, pinned out: zero-copy both ways
            else:
                # pinned PCM in, feature rows in device memory; the all-gather and the copy of the local rows to the
                # host are ordered after the kernel on the same launch lane and overlap the other lane's kernel
                with torch.cuda.stream(ext[b]):
                    if pend[b] is not None:
                        pend[b].wait()
                    if pag is not None and not fused_gather:
                        pag.wait(b, ext[b])
                    if fused_gather:
                        lsm.set_gather(pag.pointers(b), rank * B)
                    path.run_host_async(h_in, keys, out=d_feats[b], lane=b)
                    if pag is not None and not fused_gather:
                        pag.gather_async(b, d_feats[b], ext[b])
                    elif not fused_gather:
                        pend[b] = dist.all_gather_into_tensor(d_alls[b], d_feats[b], async_op=True)
                    h_feats[b].copy_(d_feats[b], non_blocking=True)

        def finish():
            for b in (0, 1):
                if pend[b] is not None:
                    with torch.cuda.stream(ext[b]):
                        pend[b].wait()
                    pend[b] = None
                if pag is not None:
                    pag.wait(b, ext[b])
            ctx.sync_all()

        for i in range(2):                 # untimed: staging buffers of this feed / sample format are allocated on first use
            one(i)
        finish()
        fence()
        t0 = time.perf_counter()
        for i in range(args.steps):
            one(i)
        finish()
        if fused_gather:
            lsm.set_gather([], 0)
        fence()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * B * args.steps / float(te.item())

    # the same steps fed with PCM16 (the samples as a WAV file stores them; converted in the kernel): half the host->device bytes
    h_pcm16 = torch.from_numpy(np.clip(np.round(pcm_np * 32768.0), -32768, 32767).astype(np.int16)).pin_memory()
    path.run_host_async(h_pcm16, keys, out=h_feats[1], lane=1)
    ctx.sync_all()
    # both ways of feeding the kernel from pinned host memory (Context.set_host_feed): the kernel's own loads over PCIe, or the
    # copy engine into a device staging buffer; and the aggregate copy-engine ceiling of the box with all ranks copying at once
    feeds = {}
    for feed in ("zero_copy", "copy_engine"):                # float32 last: h_feats are compared with the device path below
        i16_val = e2e_loop(h_pcm16, feed)
        feeds[feed] = {"float32": e2e_loop(h_pcm, feed), "pcm16": i16_val}
    ctx.set_host_feed("zero_copy")
    best_feed = max(feeds, key=lambda k: feeds[k]["float32"])
    e2e_val, e2e_i16_val = feeds[best_feed]["float32"], feeds[max(feeds, key=lambda k: feeds[k]["pcm16"])]["pcm16"]
    d_stage = torch.empty_like(d_pcm)
    fence()
    t0 = time.perf_counter()
    for _ in range(5):
        d_stage.copy_(h_pcm, non_blocking=True)
    torch.cuda.synchronize()
    h2d_s = torch.tensor([(time.perf_counter() - t0) / 5], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(h2d_s, op=dist.ReduceOp.MAX)
    h2d_ceiling_gbs = world * B * L * 4 / float(h2d_s.item()) / 1e9
    # ... and with the feature rows of the previous step going the other way at the same time, as every e2e step has it
    side = torch.cuda.Stream()
    fence()
    t0 = time.perf_counter()
    for _ in range(5):
        d_stage.copy_(h_pcm, non_blocking=True)
        with torch.cuda.stream(side):
            h_feats[1].copy_(d_feats[1], non_blocking=True)
    torch.cuda.synchronize()
    both_s = torch.tensor([(time.perf_counter() - t0) / 5], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(both_s, op=dist.ReduceOp.MAX)
    both_ceiling_gbs = world * (B * L * 4 + B * F * 8) / float(both_s.item()) / 1e9
    del d_stage
    # the same step from PAGEABLE numpy arrays through the synchronous public call (what create_dataset / extract_all_features
    # style callers hand over): the copies are staged by the driver
    e2e_pageable_val = None
    if world == 1:
        out_np = np.empty((B, F), dtype=np.float64)
        path.run_host(pcm_np, keys, out=out_np)
        t0 = time.perf_counter()
        n_pg = max(2, min(args.steps, 5))
        for _ in range(n_pg):
            path.run_host(pcm_np, keys, out=out_np)
        e2e_pageable_val = B * n_pg / (time.perf_counter() - t0)
        assert np.array_equal(out_np, d_feats[0].cpu().numpy()), "pageable host path and device path disagree"
    if pag is not None:
        # the peer-to-peer gather against NCCL's, once, outside the timed regions
        ref_all = torch.empty_like(pag.bufs[0])
        dist.all_gather_into_tensor(ref_all, d_feats[0])
        torch.cuda.synchronize()
        assert torch.equal(ref_all, pag.bufs[0]) and torch.equal(ref_all, pag.bufs[1]), "peer-to-peer all-gather differs from NCCL's"
        del ref_all
    clocks = sampler.stop() if sampler else None
    assert np.array_equal(h_feat.numpy(), d_feats[0].cpu().numpy()), "host-buffer path and device path disagree"
    if world == 1 and args.steps > 1:
        assert np.array_equal(h_feats[1].numpy(), h_feats[0].numpy()), "the two launch lanes disagree"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ctx.set_stream(None)
    fp64_peak = ctx.fp64_peak_gops()
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks["hbm_gbs"])
    step_ms = ms_total / args.steps
    fused = bool(path.fused)
    # Dominant kernel: the fused audio -> features kernel (two launches share the SMs, so the per-launch figure is taken over
    # the timed steps: kernels run back to back on the two streams for the whole region).  Binding resource:
    # the fp64 pipe (13 DFMA per channel-sample); HBM is the secondary key.
    dom_bytes = FUSED_BYTES_PER_UTT * B
    dom_gbs = dom_bytes / (step_ms / 1e3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("utterances_per_step") == B:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    spec = args.filter_mode == "speculative"
    k1_ops = K1_FP64_OPS_PER_UTT_SPEC if spec else K1_FP64_OPS_PER_UTT
    k1_gops = k1_ops * B / (k1_ms / 1e3) / 1e9
    step_gops = k1_ops * B / (step_ms / 1e3) / 1e9
    sus_step_ms = sustained_ms / n_sus
    out = {
        "metric": METRIC, "value": value, "unit": "utterances/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(),
        "neuron_steps_per_s": value * N_NEURONS * T_STEPS,
        "sustained": {"value": world * B * n_sus / (sustained_ms / 1e3), "unit": "utterances/s", "steps": n_sus, "seconds": sustained_ms / 1e3,
                      "ms_per_step": sus_step_ms, "note": "the same steps for >= 2 s, CUDA events, max over ranks"},
        "e2e": {"value": e2e_val, "unit": "utterances/s", "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * F * 8,
                "host_feed": best_feed,
                "note": "pinned host buffers through lsm_pipeline_run_host_async on alternating launch lanes, feature rows written by the "
                        "kernel into the pinned host matrix; host_feed = how the PCM reaches the kernel (Context.set_host_feed), the faster "
                        "of the two on this box - both are in e2e_feeds"},
        "e2e_feeds": {**feeds, "h2d_copy_engine_ceiling_gbs": h2d_ceiling_gbs,
                      "e2e_h2d_gbs": e2e_val * L * 4 / 1e9,
                      "h2d_plus_d2h_ceiling_gbs": both_ceiling_gbs, "e2e_h2d_plus_d2h_gbs": e2e_val * (L * 4 + F * 8) / 1e9,
                      "e2e_fraction_of_copy_ceiling": e2e_val * (L * 4 + F * 8) / 1e9 / both_ceiling_gbs,
                      "note": "utterances/s per feed and sample format; ceiling = all ranks copying their 153.6 MB batch from pinned host "
                              "memory at once with cudaMemcpyAsync (aggregate GB/s, max over ranks); h2d_plus_d2h = the same with every rank's feature "
                              "matrix going device -> pinned host concurrently (what an e2e step moves), and the e2e figure as a fraction of it"},
        "e2e_pcm16": {"value": e2e_i16_val, "unit": "utterances/s", "h2d_bytes_per_step": B * L * 2, "d2h_bytes_per_step": B * F * 8,
                      "note": "same steps with int16 PCM host buffers (lsm_pipeline_run_host_async_i16): the WAV-file sample format, "
                              "converted exactly in the kernel; e2e above keeps the float32 contract of load_audio_file"},
        "e2e_pageable": {"value": e2e_pageable_val, "unit": "utterances/s", "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * F * 8,
                         "note": "pageable numpy arrays through the synchronous lsm_pipeline_run_host (the reference-style call)"},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "roofline": {"kernel": "gammatone_encode_kernel<128,6,8,1> (fused audio->features, lane = channel: filter + encoder + reservoir + readout)"
                               if fused else "gammatone_encode_kernel (K1)",
                     "bound": "fp64", "achieved": step_gops, "peak": fp64_peak, "unit": "G fp64 lane-ops/s (DFMA = 1)",
                     "frac": step_gops / fp64_peak, "traffic": traffic,
                     "lane_ops_per_utterance": k1_ops, "launch_duration_ms": step_ms,
                     "peak_source": "measured live by lsm_fp64_peak_gops (independent DADD/DMUL register chains); nominal 148 SMs x 64 "
                                    "lanes x 1.965 GHz = 18612; MEASURED_PEAKS.json has no fp64 entry",
                     "note": "filter-bank lane-ops only (the reservoir's and the encoder's fp64 work, ~7 % more, is not counted) over the "
                             "timed steps; two launches share every SM throughout, so the step time is the per-launch time of the kernel pair",
                     "sustained_frac": k1_ops * B / (sus_step_ms / 1e3) / 1e9 / fp64_peak},
        "roofline_hbm": {"bound": "hbm", "achieved": dom_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": dom_gbs / hbm_peak,
                         "algorithmic_bytes_per_step": dom_bytes, "peak_source": f"{peak_kind} MEASURED_PEAKS.json hbm_gbs",
                         "note": "64 000 B PCM in + 51 200 B spikes + 16 000 B features out per utterance: this path is nowhere near HBM-bound"},
        "kernel_ms": {"K1_gammatone_encode": k1_ms, "K2_reservoir_features": k2_ms, "whole_step_one_caller_stream": fused_ms,
                      "K1_fp64_frac": k1_gops / fp64_peak,
                      "note": "stand-alone launches; K1 = lsm_frontend_encode (energy kernel + encoder kernel), K2 = lsm_reservoir_run, "
                              "whole step = lsm_pipeline_run back to back on one caller stream"},
        "filter_mode": {"mode": args.filter_mode, "exact_reruns_per_step": reruns_value,
                        "other_mode": other, "other_mode_K1_ms": k1_other_ms, "other_mode_step_ms": step_other_ms,
                        "other_mode_value": B / (step_other_ms / 1e3),
                        "both_modes_identical_spikes_and_features": modes_agree,
                        "note": "speculative = 13-FMA arrangement of the gammatone cascade + exact re-execution of every utterance in which "
                                "the derived distance bound could change an encoder comparison; exact = the reference's 35 separately "
                                "rounded operations per sample"},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import coracle
        cores = coracle.num_threads()
        n = min(B, max(64, cores * 24))
        sample = pcm_np[:: max(1, B // n)][:n]
        tables = host_tables()
        mask = _lib.feature_mask(keys)
        oracle_pipeline(sample[:cores], tables, lsm.reservoir, mask, 0)
        t0 = time.perf_counter()
        ref = oracle_pipeline(sample, tables, lsm.reservoir, mask, 0)
        dt = time.perf_counter() - t0
        # the two stages separately: front end + encoder, reservoir + features
        ta = time.perf_counter()
        spk_cpu = coracle.gammatone_encode(sample, *tables, [0.70, 0.80, 0.90, 0.95], 0.1, nthreads=0)
        tb = time.perf_counter()
        coracle.reservoir_run(lsm.reservoir, spk_cpu, mask, True, False, nthreads=0)
        tc = time.perf_counter()
        t1 = time.perf_counter()
        oracle_pipeline(sample[:16], tables, lsm.reservoir, mask, 1)
        dt1 = time.perf_counter() - t1
        got = path.run_host(np.ascontiguousarray(sample), keys)
        out["cpu_baseline"] = {"value": len(sample) / dt, "unit": "utterances/s", "cores": cores, "kind": "port",
                               "sample": f"{len(sample)} of the step's {B} utterances, oracle C port, {cores} threads",
                               "value_1core": 16 / dt1, "stage1_frontend_encoder": len(sample) / (tb - ta),
                               "stage23_reservoir_features": len(sample) / (tc - tb),
                               "neuron_steps_per_s": len(sample) / dt * N_NEURONS * T_STEPS,
                               "gpu_matches_cpu_bit_exact": bool(np.array_equal(got, ref))}
    if world == 1 and not os.environ.get("LSM_BENCH_NO_MEL"):
        try:
            out["mel_arm"] = mel_numbers(pcm_np, max(3, args.steps // 2), 3)
        except Exception as e:          # the mel arm is a secondary line: never lose the headline over it
            out["mel_arm"] = {"error": repr(e)}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
