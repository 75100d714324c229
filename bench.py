#!/usr/bin/env python
"""Headline benchmark: train molecules/sec of one fused fwd+bwd ELBO step (BASELINE.json metric) on the
canonical Config-B model (Conv1d encoder over one-hot 120x35, latent 292, 3x501 GRU decoder), bf16 tensor-core
mode, batch 4096 per GPU, synthetic ZINC-like ids (ZINC-250k is not in the image).

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port of the reference path

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_MOLECULE = 2.7301  # fwd+bwd, 2*MACs, layer-0 projection counted once
METRIC = "train molecules/sec (fwd+bwd ELBO)"
WORKLOAD = "Config B: models2d-stack MolecularVAE (conv enc, latent 292, 3x501 GRU, 120x35), batch 4096/GPU, bf16"
CFG = dict(latent=292, hidden=501, layers=3, seq_len=120, charset=35)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_rate(sample_b, steps, warmup):
    """molecules/sec of the numpy oracle port (fp32, BLAS threads = all host cores) on `sample_b` molecules."""
    import numpy as np
    from oracle import vae_oracle as vo
    P = vo.make_params(42, dtype=np.float32, **CFG)
    ids, onehot, eps = vo.make_batch(43, sample_b, dtype=np.float32)
    for _ in range(warmup):
        vo.config_b_step(P, onehot, eps)
    t0 = time.perf_counter()
    for _ in range(steps):
        vo.config_b_step(P, onehot, eps)
    dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps


def cpu_reference_rate(sample_b, steps, warmup):
    """(molecules/s, s/step, kind, description) of the reference's CPU path on all host cores: the UNMODIFIED reference
    modules from baseline/_ref (models2d.py:8-52 + train.py:31-38 under torch CPU, kind "reference") when the recipe
    baseline/make_ref.py has been run, else the numpy oracle port (kind "port")."""
    from baseline import ref_arm
    cores = os.cpu_count() or 1
    if ref_arm.have_ref():
        rate, sec, _ = ref_arm.step_rate(sample_b, steps, warmup, device="cpu", threads=cores)
        return rate, sec, "reference", (f"{sample_b} molecules/step, unmodified reference models2d.py + train.py:31-38 (latent 292) under "
                                        f"torch {__import__('torch').__version__} CPU fp32, torch.set_num_threads({cores})")
    rate, sec = cpu_oracle_rate(sample_b, steps, warmup)
    return rate, sec, "port", f"{sample_b} molecules/step, fp32 numpy port oracle/vae_oracle.py (baseline/_ref absent), BLAS on all cores"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    sample_b = int(max(8, min(250, 6000 // max(total, 1))))
    rate, sec, kind, sample = cpu_reference_rate(sample_b, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "molecules/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_batch": sample_b},
        "cpu_baseline": {"value": rate, "unit": "molecules/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(B):
    """The kernel to beat on the same box: the unmodified reference modules in stock PyTorch eager on
    cuda:0 (cuDNN GRU, cuBLAS, ATen) at the benchmarked batch, fp32 as shipped and under bf16 autocast."""
    import torch
    from baseline import ref_arm
    if not ref_arm.have_ref():
        return {"unavailable": "baseline/_ref absent (run baseline/make_ref.py where /root/reference exists)"}
    out = {"unit": "molecules/s", "batch": B, "impl": "reference models2d.py + train.py:31-38, torch " + torch.__version__ + " eager, cuda:0"}
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        rate, sec, loss = ref_arm.step_rate(B, 3, 2, device="cuda", autocast=ac)
        out[name] = {"value": rate, "ms_per_step": sec * 1e3, "loss": loss}
        torch.cuda.empty_cache()
    return out


def time_kernel(fn, iters=5):
    import torch
    st = torch.cuda.current_stream()
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(iters):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def kernel_rooflines(B, peaks):
    """Live CUDA-event timings of the three GEMM shapes that carry >99 % of the step's FLOPs."""
    import ctypes
    import torch
    import molecular_vae_b200 as m
    lib, vp = m._lib.lib, ctypes.c_void_p
    T, Hp = 120, 512
    st = vp(torch.cuda.current_stream().cuda_stream)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = {}

    def gemm(a, lda, amn, b, ldb, bmn, d, ldd, dbf, acc, M, N, K, bn, splits):
        m._lib.check(lib.mvae_gemm_bf16(vp(a.data_ptr()), lda, amn, vp(b.data_ptr()), ldb, bmn, vp(d.data_ptr()), ldd,
                                        dbf, acc, vp(0), M, N, K, bn, splits, vp(err.data_ptr()), st))
    x = torch.randn(T * B, Hp, device="cuda").bfloat16()
    w = (torch.randn(3 * Hp, Hp, device="cuda") * 0.04).bfloat16()
    gi = torch.empty(T * B, 3 * Hp, device="cuda", dtype=torch.bfloat16)
    t = time_kernel(lambda: gemm(x, Hp, 0, w, Hp, 0, gi, 3 * Hp, 1, 0, T * B, 3 * Hp, Hp, 256, 1))
    fl = 2.0 * T * B * 3 * Hp * Hp
    out["input_projection_gemm"] = {"ms": t * 1e3, "tflops": fl / t * 1e-12, "frac": fl / t * 1e-12 / peaks["bf16_tflops"]}
    gh = torch.empty(B, 3 * Hp, device="cuda", dtype=torch.float32)
    t = time_kernel(lambda: gemm(x, Hp, 0, w, Hp, 0, gh, 3 * Hp, 0, 0, B, 3 * Hp, Hp, 256, 1), iters=50)
    fl = 2.0 * B * 3 * Hp * Hp
    out["recurrent_step_gemm"] = {"ms": t * 1e3, "tflops": fl / t * 1e-12, "frac": fl / t * 1e-12 / peaks["bf16_tflops"]}
    dw = torch.zeros(3 * Hp, Hp, device="cuda", dtype=torch.float32)
    t = time_kernel(lambda: gemm(gi, 3 * Hp, 1, x, Hp, 1, dw, Hp, 0, 1, 3 * Hp, Hp, T * B, 256, 12))
    fl = 2.0 * T * B * 3 * Hp * Hp
    out["wgrad_gemm"] = {"ms": t * 1e3, "tflops": fl / t * 1e-12, "frac": fl / t * 1e-12 / peaks["bf16_tflops"]}
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    return out


class _BenchVocab:
    """vocab.py duck-type (30 characters + <bos>,<eos>,<pad>,<unk>, vocab.py:24) for the sampling side measurement."""

    def __init__(self, n_chars=30):
        import torch
        self.chars = [chr(ord("A") + i) for i in range(n_chars)]
        self.bos, self.eos, self.pad, self.unk = n_chars, n_chars + 1, n_chars + 2, n_chars + 3
        self.vectors = torch.eye(n_chars + 4)

    def __len__(self):
        return len(self.chars) + 4

    def string2ids(self, s, add_bos=False, add_eos=False):
        ids = [self.chars.index(c) if c in self.chars else self.unk for c in s]
        return ([self.bos] if add_bos else []) + ids + ([self.eos] if add_eos else [])

    def ids2string(self, ids, rem_bos=True, rem_eos=True):
        if ids and rem_bos and ids[0] == self.bos:
            ids = ids[1:]
        if ids and rem_eos and ids[-1] == self.eos:
            ids = ids[:-1]
        return "".join(self.chars[i] if i < len(self.chars) else "?" for i in ids)


def sampling_rate(batch=8192, max_len=100, reps=3, n_total=0, target_mean_len=45.0):
    """BASELINE.json's second metric, 'sampled SMILES/sec' (hugesample.py:113 batch 8192, mosesvae.py:214 max_len 100):
    greedy decodes of N(0,I) latents through mosesvae.VAE.sample.  `value`: device rate of the decode loop alone (CUDA events,
    ids stay on the device).  `e2e`: the reference's own definition (hugesample.py:94: strings delivered per wall second) --
    n_total latents through VAE.sample_many: prior draw, decode, device-side text assembly, device -> host copy, Python strings.
    No trained checkpoint exists (the data blobs are absent), so the weights are random-init; the <eos> bias of decoder_fc is
    calibrated so that greedy decodes end after ~target_mean_len tokens (ZINC-like) instead of never."""
    import torch
    import molecular_vae_b200 as m
    torch.manual_seed(0)
    model = m.mosesvae.VAE(_BenchVocab(), precision="bf16").cuda().eval()
    z = torch.randn(batch, 160, device="cuda")
    lo, hi = -1.0, 3.0
    for _ in range(12):          # bisection on the <eos> logit bias (monotone: larger bias -> earlier <eos>)
        mid = 0.5 * (lo + hi)
        with torch.no_grad():
            model.decoder_fc.bias[model.eos] = mid
        _, lens, _ = model.sample_ids(batch, max_len=max_len, z=z, greedy=True)
        ml = float(lens.float().mean().item())
        if ml > target_mean_len:
            lo = mid
        else:
            hi = mid
    torch.cuda.synchronize()
    model.check_device_error()
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        ids, lens, _ = model.sample_ids(batch, max_len=max_len, z=z, greedy=True)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # executed tensor-core FLOPs of one decode: per step layer 0 contracts over [onehot 64 | z 192 | h 512], layers 1-2 over [x | h],
    # each into 4 x 512 gate columns, plus the 64-column vocabulary head; (max_len - 1) steps
    flops = 2.0 * batch * (2048 * 768 + 2 * 2048 * 1024 + 64 * 512) * (max_len - 1)
    peaks, _ = load_peaks()
    out = {"metric": "sampled SMILES/sec (greedy, N(0,I) latents)", "value": batch / ms * 1e3, "unit": "SMILES/s",
           "tflops_executed": flops / (ms * 1e-3) * 1e-12,
           "frac_of_sustained_tensor_peak": flops / (ms * 1e-3) * 1e-12 / peaks["bf16_tflops_sustained"],
           "ms_per_batch": ms, "batch": batch, "max_len": max_len, "mean_len": float(lens.float().mean().item()),
           "workload": "mosesvae.VAE.sample device path (3x512 GRU decoder, d_z 160, V=34), bf16, random-init weights, <eos> bias calibrated; "
                       "persistent decode kernel (99 steps x (3 cell GEMMs + head) in one launch)"}
    if n_total > 0:
        for strs in model.sample_many(2 * batch, n_batch=batch, max_len=max_len, greedy=True, seed=1):   # warm
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n, nbytes = 0, 0
        for strs in model.sample_many(n_total, n_batch=batch, max_len=max_len, greedy=True, seed=2):
            n += len(strs)
            nbytes += len(strs[0])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["e2e"] = {"value": n / dt, "unit": "SMILES/s", "strings": n, "seconds": dt, "sample": strs[0][:60],
                      "d2h_bytes_per_batch": batch * max_len + 4 * (batch + 1), "h2d_bytes_per_batch": 8,
                      "path": "VAE.sample_many: prior draw -> 99-step decode (CUDA graph) -> ids_to_text on device -> one D2H -> list[str]"}
    return out


class _BenchCfg:   # config.py:4-85 defaults with --q_bidir (BASELINE.json configs[3]: bidirectional encoder, latent 128)
    q_cell, q_bidir, q_d_h, q_n_layers, q_dropout = "gru", True, 256, 1, 0.5
    d_cell, d_n_layers, d_dropout, d_z, d_d_h, freeze_embeddings = "gru", 3, 0, 128, 512, False


def moses_step_rate(batch=4096, reps=5, variant="mosesfile+head", rank=0, world=1):
    """BASELINE.json configs[3] at N GPUs: fused fwd+bwd step of the MOSES VAE -- mosesfile.VAE (bidirectional GRU encoder 2x256,
    latent 128, decoder GRU 3x512, V=34) with the BindingModel property head on z (logP target), or mosesvae.VAE (unidirectional,
    d_z 160, train-mode dropout 0.2) for variant "mosesvae" -- bf16, batch 4096 per GPU (SURVEY.md 8d config 4 inputs: lengths ~
    clip(N(44,9),10,98) + bos/eos, sorted descending).  One CUDA graph per phase of the step over resident inputs; with N > 1 the
    gradient bucket a phase finalises is all-reduced (NCCL AVG) while the next phase runs.  CUDA events, max over ranks by the caller."""
    import numpy as np
    import torch
    import molecular_vae_b200 as m
    torch.manual_seed(0)
    rng = np.random.Generator(np.random.PCG64(1 + rank))
    lens = np.sort(np.clip(np.rint(rng.normal(44.0, 9.0, size=batch)), 10, 98).astype(np.int64))[::-1]
    x = [torch.from_numpy(np.concatenate([[30], rng.integers(0, 30, size=int(l)), [31]]).astype(np.int64)).cuda() for l in lens]
    if variant == "mosesvae":
        model = m.mosesvae.VAE(_BenchVocab(), precision="bf16").cuda()
        target, dropout, dz = None, (0.2, 1234), 160
    else:
        model = m.mosesfile.VAE(_BenchVocab(), _BenchCfg(), precision="bf16")
        model.attach_property_head()
        model = model.cuda()
        target, dropout, dz = torch.from_numpy(rng.uniform(0, 1, size=batch).astype(np.float32)).cuda(), (0.0, 0), 128
    eps = torch.randn(batch, dz, device="cuda")
    step = m.ddp.MosesPhasedStep(model, x, eps, kl_weight=0.1, binding=target, binding_weight=1.0, dropout=dropout)
    for _ in range(2):
        step.step()
    torch.cuda.synchronize()
    model.check_device_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        step.step()
    e1.record()
    torch.cuda.synchronize()
    model.check_device_error()
    ms = e0.elapsed_time(e1) / reps
    tokens = int(sum(int(l) + 2 for l in lens))
    out = {"metric": "train molecules/sec (fwd+bwd, MOSES VAE" + (" + property head)" if target is not None else ")"),
           "value": batch / ms * 1e3, "unit": "molecules/s", "ms_per_step": ms,
           "batch": batch, "mean_len": tokens / batch, "max_len": int(lens[0]) + 2, "loss": float(model._last_scalars[0].item()),
           "workload": ("mosesfile.VAE (bidirectional encoder, d_z 128) + BindingModel head on z, joint fused step"
                        if target is not None else "mosesvae.VAE fused step (train-mode dropout 0.2)") +
                       ", packed sequences, bf16, random-init weights, batch 4096/GPU"}
    if target is not None:
        out["binding_loss"] = float(model.last_binding_loss.item())
    return out


def cfga_step_rate(batch=4096, reps=3):
    """Side measurement: the model train.py / train_distributed.py instantiate as shipped (models.py MolecularVAE, "Config A":
    Embedding + LSTM 3x72 + 3 x ConvSELU(k18) encoder, LSTM 4x1024 decoder, 32.3 M parameters) -- fused fwd + loss + bwd step,
    bf16, CUDA-graph replay over resident inputs, CUDA events.  21.39 GFLOP per molecule."""
    import numpy as np
    import torch
    import molecular_vae_b200 as m
    from oracle import vae_oracle as vo
    torch.manual_seed(42)
    model = m.models.MolecularVAE(precision="bf16").cuda()
    ids, _, eps = vo.make_batch(1, batch)
    ids, eps = torch.from_numpy(ids).cuda(), torch.from_numpy(eps).cuda()
    out = model.elbo_step(ids, eps, max_len=120)
    torch.cuda.synchronize()
    eng = model.engine(batch)
    eng.check_device_error()
    nodes = int(m._lib.lib.mvae_graph_num_kernel_nodes(eng._graph))
    eng.launch_graph()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.launch_graph()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res = {"metric": "train molecules/sec (fwd+bwd ELBO, models.py as shipped)", "value": batch / ms * 1e3, "unit": "molecules/s",
           "ms_per_step": ms, "batch": batch, "graph_nodes": nodes, "loss": float(out[0].item()),
           "tflops_algorithmic": batch / ms * 1e3 * 21.3859e-3,
           "workload": "models.py MolecularVAE (LSTM 3x72 + 3 conv encoder, LSTM 4x1024 decoder), bf16, batch 4096, per-step tcgen05 GEMMs + fused head"}
    eng.destroy_graph()
    del eng
    model._engines.clear()
    del model
    torch.cuda.empty_cache()
    return res


def rec_kernel_times(model, eng, params, ids_dev, eps_dev, steps=3):
    """Average launch duration of the persistent recurrence kernels, CUDA events on the launching stream, measured on
    DIRECT launches of the same fused step right after the timed region (a graph replay cannot carry events)."""
    import ctypes
    import torch
    import molecular_vae_b200 as m
    lib = m._lib.lib
    P, G = [p.data for p in params], [p.grad for p in params]
    eng.elbo_step(P, G, ids_dev, eps_dev)          # warm (direct path)
    torch.cuda.synchronize()
    lib.mvae_profile_enable(1)
    for _ in range(steps):
        eng.elbo_step(P, G, ids_dev, eps_dev)
    torch.cuda.synchronize()
    out = {}
    for tag, name in ((0, "fwd"), (1, "bwd")):
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        m._lib.check(lib.mvae_profile_read(tag, ctypes.byref(ms), ctypes.byref(n)))
        out[name] = (ms.value / max(n.value, 1), n.value)
    lib.mvae_profile_enable(0)
    return out


def l2_fabric_probe():
    """Measured L2 <-> SM delivery rate of this GPU (mvae_l2_probe: all SMs stream a 32 MiB L2-resident buffer with 16-byte
    L1-bypassing accesses): read-only and read+write.  The recurrence sweeps are bound by this fabric, not by the tensor pipe."""
    import ctypes
    import torch
    import molecular_vae_b200 as m
    lib, vp = m._lib.lib, ctypes.c_void_p
    nbytes, passes = 32 << 20, 24
    buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    st = vp(torch.cuda.current_stream().cuda_stream)
    out = {}
    for mode, name in ((0, "read"), (1, "copy")):
        best = 0.0
        for ctas in (148 * 2, 148 * 4):
            fn = lambda: m._lib.check(lib.mvae_l2_probe(vp(buf.data_ptr()), nbytes, passes, mode, ctas, st))
            t = time_kernel(fn, iters=5)
            moved = nbytes * passes if mode == 0 else (nbytes // 2) * passes * 2
            best = max(best, moved / t * 1e-9)
        out[name + "_gbs"] = best
    out["probe"] = "32 MiB L2-resident buffer, 24 passes, 16-byte ld.global.cg / st.global.cg from every SM, best of 296 / 592 CTAs x 512 threads"
    return out


def ncu_traffic(path, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_substr` from a committed ncu --set full summary."""
    try:
        vals, cur, take = [], None, False
        for line in open(path):
            if line.startswith("kernel:"):
                take = kernel_substr in line
                if take:
                    vals.append([0.0, 0.0])
            elif take and "dram__bytes_read.sum" in line:
                vals[-1][0] = float(line.split("=")[1].split()[0]) * (1e9 if "Gbyte" in line else 1e6 if "Mbyte" in line else 1.0)
            elif take and "dram__bytes_write.sum" in line:
                vals[-1][1] = float(line.split("=")[1].split()[0]) * (1e9 if "Gbyte" in line else 1e6 if "Mbyte" in line else 1.0)
        if not vals:
            return None
        return sum(a + b for a, b in vals) / len(vals)
    except Exception:
        return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    import molecular_vae_b200 as m
    from oracle import vae_oracle as vo  # synthetic-batch generator + cpu_baseline only
    peaks, peak_src = load_peaks()
    B = args.batch
    torch.manual_seed(42)
    model = m.VAE(precision=args.precision, **CFG).cuda()
    params = model.ordered_params()
    # one flat gradient buffer -> a single NCCL all-reduce per step in data-parallel runs
    gbuf = m.ddp.FlatGradBuffer(params)
    flat = gbuf.flat
    ids_np, _, eps_np = vo.make_batch(1000 + rank, B)
    ids_host = torch.from_numpy(ids_np).pin_memory()
    ids_dev = torch.from_numpy(ids_np).cuda()
    eps_dev = torch.from_numpy(eps_np).cuda()
    eng = model.engine(B, max_len=120)
    eng.set_train(True)
    nodes = eng.capture_elbo_step([p.data for p in params], [p.grad for p in params], ids_dev, eps_dev)
    phased, graphed = None, None
    dp_mode = "none"
    if world > 1:
        # data parallel (train_distributed.py:72): the step runs as one phase per GRU layer; the gradient bucket a phase
        # finalises is all-reduced (NCCL AVG) on NCCL's stream while the next phase's BPTT sweep runs.  Default: the phases
        # AND the all-reduces are captured into ONE CUDA graph (ddp.GraphedDataParallelStep); MVAE_DP_GRAPH=0 keeps one graph
        # per phase with the collectives issued from the host between them.
        buckets = m.ddp.phase_buckets(m.param_order(CFG["layers"]), [p.numel() for p in params], CFG["layers"])
        P_, G_ = [p.data for p in params], [p.grad for p in params]
        if os.environ.get("MVAE_DP_GRAPH", "1") != "0":
            graphed = m.ddp.GraphedDataParallelStep(lambda ph: eng.elbo_step_phase(P_, G_, ids_dev, eps_dev, ph),
                                                    CFG["layers"], flat, buckets)
            dp_mode = "one CUDA graph: 3 phases + 3 bucket all-reduces (NCCL AVG) as forked branches"
        else:
            eng.capture_elbo_step_phases(P_, G_, ids_dev, eps_dev)
            phased = m.ddp.PhasedAllReduce(flat, buckets)
            dp_mode = "3 phase graphs, bucket all-reduces (NCCL AVG) issued from the host between them"

    def step_resident():
        if graphed is not None:
            graphed.step()
            return
        if phased is None:
            eng.launch_graph()
            return
        for ph in range(CFG["layers"]):
            eng.launch_phase(ph)
            phased.after_phase(ph)
        phased.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    eng.check_device_error()
    sampler = ClockSampler(local)
    sampler.start()
    m._lib.lib.mvae_reset_launch_count()
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    for _ in range(args.steps):
        step_resident()
    e1.record(st)
    barrier()
    launches = int(m._lib.lib.mvae_launch_count())
    if graphed is not None:     # torch replays the captured graph: count what the capture enqueued, per replay
        launches = graphed.launches_per_step * args.steps
    sec = e0.elapsed_time(e1) * 1e-3
    scal = eng.scalars.cpu().numpy().tolist()

    solo_ms = None
    if world > 1:
        # every rank's own step WITHOUT the exchange (its local CUDA graph of the fused step, no barrier): the spread between
        # the GPUs of the box bounds the weak-scaling efficiency, since the collective makes all ranks wait for the slowest
        for _ in range(2):
            eng.launch_graph()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(st)
        for _ in range(5):
            eng.launch_graph()
        s1.record(st)
        torch.cuda.synchronize()
        solo = torch.tensor([s0.elapsed_time(s1) / 5], dtype=torch.float64, device="cuda")
        allsolo = [torch.zeros_like(solo) for _ in range(world)]
        dist.all_gather(allsolo, solo)
        solo_ms = [float(t.item()) for t in allsolo]
    # end to end through the public API: pinned host ids -> H2D, device-side eps draw, fused step, loss D2H
    def step_e2e():
        x = ids_host.to("cuda", non_blocking=True)
        out = model.elbo_step(x)
        if world > 1:
            dist.all_reduce(flat)
            flat.div_(world)
        return out.cpu()

    eng.destroy_graph()
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 10))
    for _ in range(e2e_steps):
        res = step_e2e()
    barrier()
    sec_e2e = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    tmax = torch.tensor([sec, sec_e2e], dtype=torch.float64, device="cuda")
    per_rank_ms = [sec / args.steps * 1e3]
    if world > 1:
        allt = [torch.zeros_like(tmax) for _ in range(world)]
        dist.all_gather(allt, tmax)
        per_rank_ms = [float(t[0].item()) / args.steps * 1e3 for t in allt]   # device time of every rank's own timed region
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    sec, sec_e2e = tmax.tolist()
    value = world * B * args.steps / sec
    e2e_value = world * B * e2e_steps / sec_e2e
    # second metric of BASELINE.json (hugesample.py): every rank decodes its own latents (replicas only, no exchange);
    # whole-job SMILES/s = batches of all ranks / slowest rank's time
    per_rank = -(-args.samples // world)
    try:
        sampling = sampling_rate(n_total=per_rank)
        if world == 1:
            # the same decode at a batch that fills whole waves of the 148 SMs: 9472 = 74 x 128 rows -> 592 = 4 x 148 work units
            # per cell GEMM (hugesample.py's 8192 gives 512 units = 3.46 waves)
            wa = sampling_rate(batch=9472, n_total=0)
            sampling["wave_aligned_batch"] = {k: wa[k] for k in ("value", "unit", "ms_per_batch", "batch", "mean_len")}
    except Exception as ex:
        sampling = {"error": repr(ex)}
    if world > 1:   # every rank takes part in the reduction, whether or not its own measurement succeeded
        t = torch.tensor([sampling.get("ms_per_batch", float("inf")), sampling.get("e2e", {}).get("seconds", float("inf"))],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if "error" not in sampling and t[0].item() != float("inf"):
            sampling["ms_per_batch"] = float(t[0].item())
            sampling["value"] = world * sampling["batch"] / sampling["ms_per_batch"] * 1e3
            sampling["replicas"] = world
            if "e2e" in sampling and t[1].item() != float("inf"):
                sampling["e2e"]["seconds"] = float(t[1].item())
                sampling["e2e"]["strings"] = world * sampling["e2e"]["strings"]
                sampling["e2e"]["value"] = sampling["e2e"]["strings"] / sampling["e2e"]["seconds"]
        elif "error" not in sampling:
            sampling = {"error": "sampling failed on another rank"}
    # dominant kernel = the BPTT sweep of one GRU layer (gru_rec2_kernel<BWD>): one launch = all T steps, timed live with CUDA
    # events on direct launches of the same step (rank 0)
    rec, rec_err = None, None
    if rank == 0:
        try:
            rec = rec_kernel_times(model, eng, params, ids_dev, eps_dev)
        except Exception as ex:
            rec_err = repr(ex)
    # BASELINE.json configs[3]: the MOSES VAE + property head step at N GPUs (every rank runs it; whole-job rate = all ranks'
    # molecules / slowest rank's time)
    eng.destroy_graph()
    eng.destroy_phase_graphs()
    del eng
    model._engines.clear()
    torch.cuda.empty_cache()
    moses_uni = None
    try:
        moses = moses_step_rate(rank=rank, world=world)
    except Exception as ex:
        moses = {"error": repr(ex)}
    if world > 1:
        t = torch.tensor([moses.get("ms_per_step", float("inf"))], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if "error" not in moses and t.item() != float("inf"):
            moses["ms_per_step"] = float(t.item())
            moses["value"] = world * moses["batch"] / moses["ms_per_step"] * 1e3
            moses["n_gpus"] = world
            moses["allreduce"] = "3 gradient buckets (readiness order), NCCL AVG, overlapped with the next phase"
        elif "error" not in moses:
            moses = {"error": "moses step failed on another rank"}
    else:
        try:
            moses_uni = moses_step_rate(variant="mosesvae")
        except Exception as ex:
            moses_uni = {"error": repr(ex)}
    if rank == 0:
        peak = peaks["bf16_tflops_sustained"]
        achieved = (value / world) * GFLOP_PER_MOLECULE * 1e-3  # TFLOP/s per GPU
        step_ms = sec / args.steps * 1e3
        # algorithmic FLOPs per launch = 2 * (3H x H) MACs per molecule-step (dh_{t-1} = dgh_t W_hh) x T x B
        H, T = CFG["hidden"], CFG["seq_len"]
        flops_launch = 2.0 * 3 * H * H * T * B
        line = {
            "metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}",
                       "allreduce": "none" if world == 1 else "3 gradient buckets (per GRU layer), overlapped with the next BPTT sweep; " + dp_mode,
                       "per_rank_ms_per_step": per_rank_ms, "per_rank_ms_per_step_without_exchange": solo_ms,
                       "cache": "per-step working set ~12 GB of activations >> 126 MB L2",
                       "graph_nodes": int(nodes), "loss": scal[0]},
            "e2e": {"value": e2e_value, "unit": "molecules/s", "h2d_bytes_per_step": int(ids_host.numel()),
                    "d2h_bytes_per_step": 16, "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
        }
        whole = {"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                 "scope": "whole fused step (algorithmic 2.7301 GFLOP/molecule x batch / step time)"}
        if rec and rec["bwd"][1] > 0:
            bwd_ms, fwd_ms = rec["bwd"][0], rec["fwd"][0]
            a = flops_launch / (bwd_ms * 1e-3) * 1e-12
            line["roofline"] = {
                "bound": "tensor", "kernel": "gru_rec2_kernel<BWD> (persistent BPTT sweep of one GRU layer, T steps per launch)",
                "achieved": a, "peak": peak, "unit": "TFLOP/s", "frac": a / peak,
                "traffic": ncu_traffic(os.path.join(ROOT, "profiles", "r02_rec2_ncu_full.txt"), "gru_rec2_kernel<1"),
                "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside back-to-back steps)",
                "flops_per_launch": flops_launch, "launch_ms": bwd_ms, "launches_timed": rec["bwd"][1],
                "share_of_step": CFG["layers"] * bwd_ms / step_ms,
                "fwd_sweep": {"launch_ms": fwd_ms, "achieved": flops_launch / (fwd_ms * 1e-3) * 1e-12,
                              "frac": flops_launch / (fwd_ms * 1e-3) * 1e-12 / peak,
                              "share_of_step": CFG["layers"] * fwd_ms / step_ms},
                "whole_step": whole,
            }
        else:
            line["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                                "frac": achieved / peak, "traffic": None, "peak_source": peak_src + " bf16_tflops_sustained",
                                "scope": whole["scope"], "kernel_timing_error": rec_err}
        if world == 1 and rec and rec["bwd"][1] > 0:
            try:
                # what actually bounds the sweeps: bytes moved between L2 and the SMs per launch (streamed operand
                # re-read by the unit slices + saved gates + dX / gi in, dG / hs / saved gates out) against the measured fabric rate
                Bp, Hp = -(-B // 256) * 256, -(-H // 64) * 64
                slab = Bp * Hp * 2.0
                bwd_bytes = T * (4 * 3 * slab + 5 * slab + slab + 4 * slab)          # 4 unit-slice clusters x dgh, sv, dX, dG
                fwd_bytes = T * (8 * slab + 3 * slab + 5 * slab + slab)              # 8 unit slices x h, gi, sv, hs
                l2 = l2_fabric_probe()
                l2["bwd_sweep"] = {"l2_bytes_per_launch": bwd_bytes, "achieved_gbs": bwd_bytes / (rec["bwd"][0] * 1e-3) * 1e-9}
                l2["fwd_sweep"] = {"l2_bytes_per_launch": fwd_bytes, "achieved_gbs": fwd_bytes / (rec["fwd"][0] * 1e-3) * 1e-9}
                for k in ("bwd_sweep", "fwd_sweep"):
                    l2[k]["frac_of_copy_probe"] = l2[k]["achieved_gbs"] / l2["copy_gbs"]
                    l2[k]["frac_of_read_probe"] = l2[k]["achieved_gbs"] / l2["read_gbs"]
                line["roofline"]["l2_fabric"] = l2
            except Exception as ex:
                line["roofline"]["l2_fabric_error"] = repr(ex)
        if world == 1:
            try:
                line["roofline"]["kernels"] = kernel_rooflines(B, peaks)
            except Exception as ex:  # never lose the headline over the side measurement
                line["roofline"]["kernels_error"] = repr(ex)
            cores = os.cpu_count() or 1
            rate, _, kind, sample = cpu_reference_rate(250, 4, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "molecules/s", "cores": cores, "kind": kind,
                                    "sample": "4 timed steps after 1 warm-up (BASELINE config[0] batch): " + sample}
            try:
                line["gpu_eager_baseline"] = gpu_eager_baseline(B)
            except Exception as ex:
                line["gpu_eager_baseline"] = {"error": repr(ex)}
        line["sampling"] = sampling
        line["moses_step"] = moses
        if world == 1:
            try:
                line["cfga_step"] = cfga_step_rate()
            except Exception as ex:
                line["cfga_step"] = {"error": repr(ex)}
        if moses_uni is not None:
            line["moses_step_unidirectional"] = moses_uni
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that hold captured NCCL kernels must be gone before the communicator is torn down; the tear-down itself
        # (ncclCommDestroy behind destroy_process_group) was seen to hang after such graphs on this image, so every rank leaves
        # through a final barrier + _exit once its line is out (exit status 0, nothing left to flush)
        graphed = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=10_000_000,
                    help="latents decoded end to end for the sampling metric (BASELINE.json configs[4]: 10M), split over the ranks")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
