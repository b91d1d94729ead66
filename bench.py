#!/usr/bin/env python
"""bench.py — QA-examples/s of the STAGE forward hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one eval forward (`--add_local`, both streams) over one synthetic batch of BASELINE configs[1]
(batch 32, T=300 frames, R=20 regions, Lq=Ls=100 tokens, H=128, dense masks = the throughput shape of SURVEY.md §8d).
Prints ONE JSON line (rank 0):
  value     examples/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the public module call with HOST (pinned) inputs: every step uploads its inputs
            host->device and reads answer + span probabilities back, copies inside the timed region (double-buffered
            on a copy stream)
  roofline  the dominant kernel (cqa_kernel) timed live on its own stream: algorithmic FLOPs / duration vs the measured
            tensor peak in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle port (same torch ops as the reference) on the host cores, bounded sample
  eager_cuda_baseline / parity_at_bench_config  the reference's eager path on this GPU; the timed batch checked against the oracle
  train_step    (N=1, default config) the line of `bench.py --config cfg3` - BASELINE configs[2], the full training step - run
                as a child process with a 300 s timeout
`--config cfg5` / `--heads 4` / `--config cfg3` select the other BASELINE configurations (long context, self-attention
encoders, training step: forward + backward + clip + Adam, one NCCL gradient all-reduce per step when N > 1).
`--impl reference` times only that CPU port (all host threads), one example per step.
Inputs (3.3 GB / batch) are far larger than the 126 MB L2, so no explicit L2 flush is needed between iterations.
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

import torch  # noqa: E402

METRIC = "QA-examples/sec forward (synthetic T=300,R=20,L=100,H=128)"
SHAPE = dict(T=300, R=20, Lq=100, Ls=100)
# BASELINE.json configs: cfg2 = configs[1] (the headline line), cfg5 = configs[4] (long-context stress)
CONFIGS = {"cfg2": dict(T=300, R=20, Lq=100, Ls=100, batch=32, name="configs[1]"),
           "cfg5": dict(T=600, R=40, Lq=100, Ls=100, batch=8, name="configs[4] long-context"),
           "cfg1": dict(T=40, R=25, Lq=40, Ls=50, batch=2, name="configs[0] debug shapes"),
           # configs[2] / configs[3]: full train step (forward + backward + clip + Adam [+ NCCL gradient all-reduce when N > 1])
           "cfg3": dict(T=300, R=20, Lq=100, Ls=100, batch=16, name="configs[2] train step", train=True)}
H = 128


def alg_flops_cqa(n_ex, T, R, Lq, Ls):
    """Algorithmic FLOPs of a7+a8+a9: S and U for both streams + three 384->128 projections."""
    n_st = 5 * T * Lq
    return n_ex * (4 * H * n_st * (Ls + R) + 3 * (2 * 3 * H * H) * n_st)


def alg_flops_total(n_ex, T, R, Lq, Ls, k_in=7, k_cls=5, n_conv=2, heads=0):
    n_qa, n_sub, n_vid, n_st = 5 * Lq, T * Ls, T * R, 5 * T * Lq
    f = 2 * 768 * 300 * (n_qa + n_sub) + 2 * 300 * 300 * n_vid + 2 * 300 * H * (n_qa + n_sub + n_vid)
    f += n_conv * (2 * k_in * H + 2 * H * H) * (n_qa + n_sub + n_vid)
    f += 4 * H * n_st * (Ls + R) + 3 * (2 * 3 * H * H) * n_st + n_conv * (2 * k_cls * H + 2 * H * H) * n_st
    f += (2 * H * H + 4 * H) * 5 * T
    if heads:  # 8H^2 per token + 4H*L per token (QK^T and PV) for each encoder
        f += 8 * H * H * (n_qa + n_sub + n_vid + n_st) + 4 * H * (n_qa * Lq + n_sub * Ls + n_vid * R + n_st * Lq)
    return n_ex * f


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d.get("bf16_tflops_sustained", 1416.1)), hbm=float(d.get("hbm_gbs", 6458.1)), src="measured")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def make_host_batch(N, pin, shape):
    """Dense synthetic batch of the chosen BASELINE config on the host (pinned for the e2e leg)."""
    from tvqaplus_b200 import host
    return host.synthetic_batch(N, shape["T"], shape["R"], shape["Lq"], shape["Ls"], seed=2018, pin=pin)


TENSOR_KEYS = ("qas_bert", "sub_bert", "vid", "qas_mask", "sub_mask", "vid_mask", "ts_label_mask", "target")


def to_device(b, dev, non_blocking=False, into=None):
    from tvqaplus_b200 import host
    out = host.Batch(b)
    for k in TENSOR_KEYS:
        if into is not None:
            into[k].copy_(b[k], non_blocking=non_blocking)
            out[k] = into[k]
        else:
            out[k] = b[k].to(dev, non_blocking=non_blocking)
    out["ts_label"] = {k: v.to(dev) for k, v in b["ts_label"].items()} if into is None else into["ts_label"]
    return out


def take_examples(b, idx):
    """Examples ``idx`` of a host batch as a batch of their own (the CPU legs run on a bounded sample of the timed batch)."""
    from tvqaplus_b200 import host
    out = host.Batch(b)
    n = len(b["qid"])
    for k, v in b.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
            out[k] = v[idx].contiguous()
        elif isinstance(v, list) and len(v) == n:
            out[k] = [v[i] for i in idx]
    out["ts_label"] = {k: v[idx] for k, v in b["ts_label"].items()}
    return out


def cpu_port(batch, opt, sd, steps, warmup, threads, device="cpu"):
    """Times the oracle port (torch fp32 ops in the reference's order: the reference's own path restated) on ``device``.
    Returns (times, last outputs)."""
    from oracle import stage_oracle as O
    torch.set_num_threads(threads)
    times, out = [], None
    for i in range(warmup + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = O.stage_forward(sd, opt, batch, keep=True)
        if device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, out


def cpu_train_port(batch, opt, sd, steps, warmup, threads, device="cpu"):
    """One training step of the reference's loop (main.py:53-66) on the oracle port with torch autograd: forward, the three
    losses, backward, clip_grad_norm_(10), Adam(lr 1e-3, wd 3e-7).  Returns the step times."""
    import torch.nn.functional as F
    from oracle import stage_oracle as O
    torch.set_num_threads(threads)
    if getattr(opt, "dropout", 0) > 0:  # torch's own dropout at the reference's 21 sites (the oracle's hook)
        O.DROP = lambda tag, x: F.dropout(x, opt.dropout, True)
    ref = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(".pe") else v.clone()) for k, v in sd.items()}
    params = [v for v in ref.values() if v.requires_grad]
    optim = torch.optim.Adam(params, lr=1e-3, weight_decay=3e-7)
    n = len(batch["qid"])
    times = []
    for i in range(warmup + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.enable_grad():
            out = O.stage_forward.__wrapped__(ref, opt, batch, training=True)
            loss = F.cross_entropy(out["answer"], out["target"], reduction="sum") * (1.0 * n / len(out["target"])) + 0.5 * out["temporal_loss"]
            if batch.get("att_labels") is not None:
                loss = loss + 0.1 * O.get_att_loss(out["vid_raw_s"], batch["att_labels"], batch["target"], batch["anno_st_idx"],
                                                   opt.num_negatives, opt.att_loss_type, opt.alpha, opt.margin)
        optim.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 10.0)
        optim.step()
        if device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    O.DROP = None
    return times


def train_main(args, cfg, rank, local_rank, world, cores):
    """--config cfg3: examples/s of the full training step (BASELINE.json configs[2]; configs[3] when launched on N > 1 GPUs:
    one NCCL SUM all-reduce of the flat gradient buffer per step)."""
    from tvqaplus_b200 import host
    from tvqaplus_b200.model import STAGE
    shape = {k: cfg[k] for k in ("T", "R", "Lq", "Ls")}
    T, R, Lq, Ls = shape["T"], shape["R"], shape["Lq"], shape["Ls"]
    N = args.batch or cfg["batch"]
    metric = "QA-examples/sec full train step (forward+backward+clip+Adam%s; synthetic T=%d,R=%d,L=%d,H=128)" % (
        "+NCCL grad all-reduce" if world > 1 else "", T, R, Lq)
    opt = host.default_opt(dropout=args.dropout)
    config = dict(workload="%s: train mode, add_local, use_sup_att, sub+vfeat, heads=0, dropout=%g (config.py:105 default 0.1), "
                           "batch=%d/GPU, T=%d, R=%d, Lq=%d, Ls=%d, H=128, dense masks; fp32 throughout (the reference's dtype; BASELINE says bf16)"
                  % (cfg["name"] if world == 1 else "configs[3] data-parallel train step", args.dropout, N, T, R, Lq, Ls),
                  global_batch=N * world, per_gpu_batch=N,
                  parallelism="dp%d (one process per GPU, %s)" % (world, "NCCL SUM all-reduce of the flat 2.2 MB gradient buffer" if world > 1 else "no collective"),
                  l2="activations (tens of GB per step) exceed the 126 MB L2; no flush needed")

    if args.impl == "reference":
        if rank != 0:
            return
        W = max(1, min(args.warmup, 1))
        torch.manual_seed(2018)
        sd = STAGE(opt).state_dict()
        one = take_examples(make_host_batch(2, False, shape), [0])
        one["att_labels"], one["anno_st_idx"] = host.synthetic_att_labels(one, seed=1)
        times = cpu_train_port(one, opt, sd, args.steps, W, cores)
        v = 1.0 / statistics.mean(times)
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "examples/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": W, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": dict(config, sample="1 example of the shape per step"),
                          "cpu_baseline": {"value": v, "unit": "examples/s", "cores": cores, "kind": "port",
                                           "sample": "CPU oracle port + torch autograd + clip + Adam, 1 example/step, %d steps" % args.steps},
                          "e2e": {"value": v, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    from tvqaplus_b200 import _lib, ops
    from tvqaplus_b200.train import Trainer
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the STAGE kernels)")
    if world > 1:
        import datetime
        import torch.distributed as dist
        # a short collective timeout: a rank that stops answering aborts the run in minutes instead of holding 8 GPUs
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.lib()
    torch.manual_seed(2018)
    model = STAGE(opt).to(dev).train()
    trainer = Trainer(model)
    hbatch = make_host_batch(N, pin=True, shape=shape)
    hbatch["target"] = (hbatch["target"] + rank) % 5  # ranks see different examples
    hbatch["att_labels"], hbatch["anno_st_idx"] = host.synthetic_att_labels(hbatch, seed=1 + rank)
    dbatch = to_device(hbatch, dev)
    for k in ("att_labels", "anno_st_idx"):
        dbatch[k] = hbatch[k]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # CUDA-graph replay of the step: verified on one GPU.  With N > 1 the collective sits between two graphs (train.py); that
    # path has never run on hardware (the round's GPU budget ended on the first 8-GPU attempt), so it is opt-in there.
    use_graph = not args.no_graph and (world == 1 or args.graph_multi)

    def one_step(b):
        # the sampled (positive, negative) pairs are drawn on the host every step, like the reference (model/stage.py:613-694)
        pos, neg = model._att_pairs(b["att_labels"], hbatch["target"], b["anno_st_idx"], model.num_negatives, False, model.drop_topk)
        return trainer.step(b, trainer.engine.att_offsets_from_pairs(pos, neg, N, T, Lq, R), graph=use_graph)

    from tvqaplus_b200 import train_ops
    losses = []
    ops.LAUNCHES = train_ops.LAUNCHES = 0
    losses.append(float(one_step(dbatch)["loss"]))  # eager, and (with graphs) the capture of the step for the replays below
    per_step_launches = (ops.LAUNCHES + train_ops.LAUNCHES) // (2 if use_graph else 1)
    for _ in range(args.warmup - 1):
        losses.append(float(one_step(dbatch)["loss"]))
    barrier()
    peak0 = torch.cuda.max_memory_allocated(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for _ in range(args.steps):
            last = one_step(dbatch)["loss"]
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = N * world * args.steps / (ms_total * 1e-3)
    losses.append(float(last))

    # end to end: the batch travels from pinned host memory every step, the loss is read back every step
    e2e = None
    if not args.no_e2e:
        copy_stream, main_stream = torch.cuda.Stream(), torch.cuda.current_stream()
        bufs = [{k: torch.empty_like(dbatch[k]) for k in TENSOR_KEYS} for _ in range(2)]
        for b in bufs:
            b["ts_label"] = dbatch["ts_label"]
        ready, freed = [torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.empty(1).pin_memory()
        h2d = sum(hbatch[k].numel() * hbatch[k].element_size() for k in TENSOR_KEYS)

        def run_e2e(steps):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(copy_stream):
                s0.record()
                to_device(hbatch, dev, non_blocking=True, into=bufs[0])
                ready[0].record()
            for i in range(steps):
                cur = i % 2
                if i + 1 < steps:
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(freed[1 - cur])
                        to_device(hbatch, dev, non_blocking=True, into=bufs[1 - cur])
                        ready[1 - cur].record()
                main_stream.wait_event(ready[cur])
                b = host.Batch(hbatch)
                b.update(bufs[cur])
                res = one_step(b)
                loss_host.copy_(res["loss"].reshape(1), non_blocking=True)
                freed[cur].record()
            s1.record()
            torch.cuda.synchronize()
            return s0.elapsed_time(s1)

        run_e2e(2)
        barrier()
        ms_e2e = max_over_ranks(run_e2e(args.steps))
        barrier()
        e2e = {"value": N * world * args.steps / (ms_e2e * 1e-3), "unit": "examples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e / args.steps,
               "note": "pinned host inputs, double-buffered H2D on a copy stream, the loss read back every step"}
        del bufs

    pk = peaks()
    flops = 3 * alg_flops_total(N, T, R, Lq, Ls)  # forward + two gradient products per contraction
    achieved = flops * args.steps / (ms_total * 1e-3) / 1e12 if world == 1 else flops * world * args.steps / (ms_total * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "sgemm_kernel (all GEMMs of the step)", "achieved": achieved / world, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / world / pk["tflops"], "traffic": None, "peak_source": pk["src"] + " bf16 sustained",
                "note": "per GPU; algorithmic FLOPs of the step = 3 x forward (%.1f GFLOP/example); the training GEMMs run in exact fp32 "
                        "on the CUDA cores (FFMA peak of a B200 is ~75 TFLOP/s), not on the tensor cores" % (flops / N / 1e9)}
    cpu = eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        one = take_examples(hbatch, [0])
        one["att_labels"], one["anno_st_idx"] = [hbatch["att_labels"][0]], [hbatch["anno_st_idx"][0]]
        times = cpu_train_port(one, opt, sd_cpu, 2, 1, cores)
        cpu = {"value": 1.0 / statistics.mean(times), "unit": "examples/s", "cores": cores, "kind": "port",
               "sample": "CPU oracle port + torch autograd + clip + Adam on example 0 of the timed batch, 1 warm-up + 2 timed steps"}
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        try:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            sd_dev = {k: v.detach().clone() for k, v in model.state_dict().items()}
            del trainer
            torch.cuda.empty_cache()
            nb = min(N, 8)
            sub = take_examples(hbatch, list(range(nb)))
            sub["att_labels"], sub["anno_st_idx"] = hbatch["att_labels"][:nb], hbatch["anno_st_idx"][:nb]
            sub = to_device(sub, dev)
            sub["att_labels"], sub["anno_st_idx"] = hbatch["att_labels"][:nb], hbatch["anno_st_idx"][:nb]
            times = cpu_train_port(sub, opt, sd_dev, 2, 1, cores, device="cuda")
            eager = {"value": nb / statistics.mean(times), "unit": "examples/s", "ms_per_step": 1e3 * statistics.mean(times), "batch": nb, "kind": "port",
                     "note": "the oracle port with torch autograd, clip_grad_norm_ and torch.optim.Adam, eager on this GPU, fp32 with "
                             "allow_tf32=False, batch %d (autograd keeps every intermediate: batch 16 needs ~93 GB), 1 warm-up + 2 timed" % nb}
        except Exception as e:
            eager = {"unavailable": "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:200])}
    if rank == 0:
        print(json.dumps({"metric": metric, "value": value, "unit": "examples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config, "clocks": clk.summary(), "e2e": e2e,
                          "gpu_launches": per_step_launches * args.steps,
                          "launch_mode": "eager (one Python call per kernel)" if not use_graph else "the whole step replayed from one CUDA graph (%d kernel launches per step)" % per_step_launches,
                          "roofline": roofline, "cpu_baseline": cpu, "eager_cuda_baseline": eager,
                          "loss_first_last": [losses[0], losses[-1]], "peak_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json config (cfg2 = configs[1], the headline)")
    ap.add_argument("--heads", type=int, default=0, choices=[0, 4], help="self-attention heads in both encoders (config.py:110,116)")
    ap.add_argument("--batch", type=int, default=0, help="examples per GPU per step (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the configs[2] training-step line that rides along as `train_step`")
    ap.add_argument("--dropout", type=float, default=0.1, help="train configs: dropout probability (config.py:105)")
    ap.add_argument("--graph-multi", action="store_true", help="train configs on N > 1 GPUs: replay two CUDA graphs around the eager all-reduce (untested)")
    ap.add_argument("--no-graph", action="store_true", help="train configs: launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    cfg = CONFIGS[args.config]
    if cfg.get("train"):
        return train_main(args, cfg, rank, local_rank, world, cores)
    shape = {k: cfg[k] for k in ("T", "R", "Lq", "Ls")}
    T, R, Lq, Ls = shape["T"], shape["R"], shape["Lq"], shape["Ls"]
    N = args.batch or cfg["batch"]
    in_gb = N * 4 * (5 * Lq * 768 + T * Ls * 768 + T * R * 300) / 1e9
    config = dict(workload="%s: eval forward, add_local, sub+vfeat, heads=%d, batch=%d/GPU, T=%d, R=%d, Lq=%d, Ls=%d, H=128, dense masks"
                  % (cfg["name"], args.heads, N, T, R, Lq, Ls), global_batch=N * world, per_gpu_batch=N,
                  parallelism="dp%d (replicas, no collective)" % world,
                  l2="inputs (%.2f GB/batch) %s the 126 MB L2%s" % (in_gb, "exceed" if in_gb > 0.2 else "FIT in", "; no flush needed" if in_gb > 0.2 else " (numerics-gate shape, not a timing shape)"))
    metric = METRIC if args.config == "cfg2" else "QA-examples/sec forward (synthetic T=%d,R=%d,L=%d,H=128)" % (T, R, Lq)

    from tvqaplus_b200 import host
    from tvqaplus_b200.model import STAGE
    opt = host.default_opt(input_encoder_n_heads=args.heads, cls_encoder_n_heads=args.heads)

    if args.impl == "reference":
        if rank != 0:
            return
        W = max(1, min(args.warmup, 2))
        torch.manual_seed(2018)
        sd = STAGE(opt).state_dict()
        one = take_examples(make_host_batch(2, False, shape), [0])
        times, _ = cpu_port(one, opt, sd, args.steps, W, cores)
        v = 1.0 / statistics.mean(times)
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "examples/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": W, "ms_per_step": 1e3 * statistics.mean(times),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": dict(config, sample="1 example of the %s shape per step" % cfg["name"]),
                          "cpu_baseline": {"value": v, "unit": "examples/s", "cores": cores, "kind": "port",
                                           "sample": "CPU oracle port (torch fp32, reference op order), 1 example/step, %d steps" % args.steps},
                          "e2e": {"value": v, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    from tvqaplus_b200 import _lib, ops
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the STAGE kernels)")
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.lib()
    torch.manual_seed(2018)
    model = STAGE(opt).to(dev).eval()
    hbatch = make_host_batch(N, pin=not args.no_e2e, shape=shape)
    dbatch = to_device(hbatch, dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- device-resident throughput
    for _ in range(args.warmup):
        model(dbatch)
    ops.LAUNCHES = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for _ in range(args.steps):
            out = model(dbatch)
        e1.record()
        barrier()
    launches = ops.LAUNCHES
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    _lib.check_device_error()
    value = N * world * args.steps / (ms_total * 1e-3)
    dev_answer, dev_probs, dev_spans = out[0].cpu(), out[4].cpu(), model._last_spans.cpu().long().view(N, 5, 2)

    # ---------------------------------------------------------------- end-to-end (host buffers, copies inside)
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        main_stream = torch.cuda.current_stream()
        bufs = [{k: torch.empty_like(dbatch[k]) for k in TENSOR_KEYS} for _ in range(2)]
        for b in bufs:
            b["ts_label"] = dbatch["ts_label"]
        res_host = [dict(a=torch.empty(N, 5).pin_memory(), p=torch.empty(N, 5, T, 2).pin_memory()) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        h2d = sum(hbatch[k].numel() * hbatch[k].element_size() for k in TENSOR_KEYS)
        d2h = res_host[0]["a"].numel() * 4 + res_host[0]["p"].numel() * 4

        def run_e2e(steps):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(copy_stream):
                s0.record()
                to_device(hbatch, dev, non_blocking=True, into=bufs[0])
                ready[0].record()
            for i in range(steps):
                cur = i % 2
                if i + 1 < steps:
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(freed[1 - cur])
                        to_device(hbatch, dev, non_blocking=True, into=bufs[1 - cur])
                        ready[1 - cur].record()
                main_stream.wait_event(ready[cur])
                b = host.Batch(hbatch)
                b.update(bufs[cur])
                ans, _, _, _, probs = model(b)
                res_host[cur]["a"].copy_(ans, non_blocking=True)
                res_host[cur]["p"].copy_(probs, non_blocking=True)
                freed[cur].record()
            s1.record()
            torch.cuda.synchronize()
            return s0.elapsed_time(s1)

        run_e2e(2)
        barrier()
        ms_e2e = max_over_ranks(run_e2e(args.steps))
        barrier()
        e2e = {"value": N * world * args.steps / (ms_e2e * 1e-3), "unit": "examples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
               "h2d_gb_per_s_per_rank": h2d / (ms_e2e / args.steps * 1e-3) / 1e9,
               "note": "pinned host inputs, double-buffered H2D on a copy stream, answer + span probs read back every step"}
        del bufs

        # ------------------------------------------------------------ end-to-end on the RAGGED variant of the same config
        # (QA length ~U[0.5 Lq, Lq], sub-title tokens ~U[0.3 Ls, Ls], regions ~U[0.5 R, R], frames ~U[0.7 T, T]):
        # the collate keeps only the valid rows (tvqaplus_b200/collate.py), every step uploads them from pinned memory and pads
        # on the device (stage_pad_rows) before the forward; answer + span probabilities are read back as above.
        from tvqaplus_b200 import collate
        rb = collate.pad_collate(collate.synthetic_items(N, T, R, Lq, Ls, seed=2018 + rank), pin=True)
        maxl = dict(max_qa_l=Lq, max_vid_l=T, max_sub_l=Ls)
        rg_h2d = rb.h2d_bytes()

        def run_ragged(steps):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            staged, evs = [None, None], [torch.cuda.Event(), torch.cuda.Event()]

            def upload(slot):
                with torch.cuda.stream(copy_stream):
                    inp, _, _ = collate.prepare_inputs(rb, maxl, dev)
                    for v in inp.values():
                        if torch.is_tensor(v):
                            v.record_stream(main_stream)
                    staged[slot] = inp
                    evs[slot].record()

            with torch.cuda.stream(copy_stream):
                s0.record()
            upload(0)
            for i in range(steps):
                cur = i % 2
                if i + 1 < steps:
                    upload(1 - cur)
                main_stream.wait_event(evs[cur])
                ans, _, _, _, probs = model(staged[cur])
                res_host[cur]["a"].copy_(ans, non_blocking=True)
                res_host[cur]["p"][:, :, :probs.shape[2]].copy_(probs, non_blocking=True)
            s1.record()
            torch.cuda.synchronize()
            return s0.elapsed_time(s1)

        run_ragged(2)
        barrier()
        ms_rg = max_over_ranks(run_ragged(args.steps))
        barrier()
        e2e["ragged"] = {"value": N * world * args.steps / (ms_rg * 1e-3), "unit": "examples/s", "h2d_bytes_per_step": rg_h2d,
                         "ms_per_step": ms_rg / args.steps, "h2d_gb_per_s_per_rank": rg_h2d / (ms_rg / args.steps * 1e-3) / 1e9,
                         "bytes_vs_dense": rg_h2d / h2d,
                         "note": "ragged masks of SURVEY 8d; valid rows only travel host->device, padding + masks are built on the device "
                                 "(tvqaplus_b200.collate.prepare_inputs = the reference's prepare_inputs signature)"}
        del rb

    # ---------------------------------------------------------------- roofline of the dominant kernel (rank 0 timing, own stream)
    with torch.no_grad():
        a_embed = model.base_encoder(dbatch.qas_bert.view(N * 5, Lq, 768), dbatch.qas_mask.view(N * 5, Lq), model._text_fc, None, model.input_encoder).view(N, 5, Lq, H)
        vid_embed = model.base_encoder(dbatch.vid.view(N * T, R, 300), dbatch.vid_mask.view(N * T, R), model._vid_fc, None, model.input_encoder).view(N, T, R, H)
        # like STAGE.forward: the sub-title encoder writes its output into the context blobs of the CQA workspace, so the
        # timed launch is what the forward runs (preparation of answers + regions, a.Wa, the fused kernel)
        ws = ops.CqaWorkspace(N, T, Lq, dev)
        sub_embed = None
        if model.input_encoder.can_write_cqa_workspace(Ls):
            model.input_encoder(model._text_fc(dbatch.sub_bert.view(N * T, Ls, 768)), dbatch.sub_mask.view(N * T, Ls), cqa_workspace=ws)
        else:
            sub_embed = model.base_encoder(dbatch.sub_bert.view(N * T, Ls, 768), dbatch.sub_mask.view(N * T, Ls), model._text_fc, None, model.input_encoder).view(N, T, Ls, H)
        w = model._cqa_weights()
        blk = ops.blocked_fits(Lq)  # the same statement layout STAGE.forward uses
        run_cqa = lambda: ops.cqa_fused_fwd(a_embed, dbatch.qas_mask, sub_embed, dbatch.sub_mask, vid_embed, dbatch.vid_mask, w, opt.scale,
                                            blocked=blk, workspace=ws)
        run_cqa()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        k0.record()
        for _ in range(reps):
            run_cqa()
        k1.record()
        torch.cuda.synchronize()
        del ws, a_embed, vid_embed, sub_embed
    cqa_ms = k0.elapsed_time(k1) / reps
    pk = peaks()
    achieved = alg_flops_cqa(N, T, R, Lq, Ls) / (cqa_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "cqa_traffic.json")
    if os.path.exists(tp) and args.config == "cfg2" and N == 32:
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    roofline = {"bound": "tensor", "kernel": "cqa_kernel", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": traffic, "traffic_source": traffic_src, "ms_per_launch": cqa_ms,
                "peak_source": pk["src"] + " bf16 sustained",
                "note": "algorithmic FLOPs (a7+a8+a9, %.1f GFLOP/example); executed tensor FLOPs are 3x (split-fp16) on 128-row tiles"
                        % (alg_flops_cqa(1, T, R, Lq, Ls) / 1e9),
                "step_share": cqa_ms / (ms_total / args.steps)}

    # ---------------------------------------------------------------- CPU baseline + parity at the benchmarked configuration
    # (rank 0, N=1 only): the oracle port runs on two examples OF THE TIMED BATCH (the first and the last), which times the
    # reference's CPU path and checks the outputs the timed forward produced for exactly those rows.
    cpu = parity = eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        idx = [0, N - 1] if N > 1 else [0]
        sample = take_examples(hbatch, idx)
        sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        times, want = cpu_port(sample, opt, sd_cpu, 3, 1, cores)
        cpu = {"value": len(idx) / statistics.mean(times), "unit": "examples/s", "cores": cores, "kind": "port",
               "sample": "CPU oracle port (torch fp32, reference op order) on examples %s of the timed batch, 1 warm-up + 3 timed" % idx}

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max())

        parity = {"examples": idx, "of_batch": N,
                  "answer_max_rel": rel(dev_answer[idx], want["answer"]), "t_probs_max_rel": rel(dev_probs[idx], want["t_probs"]),
                  "answer_argmax_equal": bool(torch.equal(dev_answer[idx].argmax(1), want["answer"].argmax(1))),
                  "spans_equal": bool(torch.equal(dev_spans[idx].reshape(-1, 2), want["spans"])),
                  "tolerance": 1e-3, "checker": "oracle/stage_oracle.py on the host, fp32"}
    # ---------------------------------------------------------------- the honest competitor: the reference's eager path on this GPU
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        try:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            sd_dev = {k: v.detach() for k, v in model.state_dict().items()}
            torch.cuda.empty_cache()
            times, _ = cpu_port(dbatch, opt, sd_dev, 3, 1, cores, device="cuda")
            eager = {"value": N / statistics.mean(times), "unit": "examples/s", "ms_per_step": 1e3 * statistics.mean(times),
                     "batch": N, "kind": "port",
                     "note": "the oracle port = the reference's torch ops in the reference's order, eager on this GPU, fp32 with "
                             "allow_tf32=False, same batch, wall clock with synchronize (includes the reference's host-side numpy "
                             "span search, model/model_utils.py:102-121); 1 warm-up + 3 timed"}
        except Exception as e:  # e.g. out of memory at a large config: report, never fail the bench line
            eager = {"unavailable": "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:200])}

    # ---------------------------------------------------------------- the training step (BASELINE configs[2]) beside the headline
    # `bench.py --config cfg3` run as a child process with a hard timeout (its own line, minus the slow baselines), so that a
    # problem there can never cost the headline line; rank 0 at N=1 on the default configuration only.
    train_step = None
    if rank == 0 and world == 1 and args.config == "cfg2" and args.heads == 0 and not args.no_train_step:
        try:
            del dbatch, hbatch
            torch.cuda.empty_cache()
            cmd = [sys.executable, os.path.abspath(__file__), "--config", "cfg3", "--steps", "5", "--warmup", "3",
                   "--no-cpu-baseline", "--no-eager-baseline", "--no-train-step"]
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
            rows = [l for l in res.stdout.splitlines() if l.startswith("{")]
            if res.returncode == 0 and rows:
                t = json.loads(rows[-1])
                train_step = {k: t.get(k) for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "config", "e2e", "launch_mode",
                                                    "gpu_launches", "loss_first_last", "peak_memory_gb", "clocks")}
            else:
                train_step = {"unavailable": "exit code %d: %s" % (res.returncode, (res.stderr.strip().splitlines() or ["no output"])[-1][:200])}
        except subprocess.TimeoutExpired:
            train_step = {"unavailable": "timed out after 300 s"}
        except Exception as e:
            train_step = {"unavailable": "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:200])}

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "examples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 (split-fp16 x3 tensor-core GEMMs, fp32 accumulate)", "data": "synthetic",
                "config": config, "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu, "parity_at_bench_config": parity, "eager_cuda_baseline": eager,
                "alg_tflops_whole_path": alg_flops_total(N * world, T, R, Lq, Ls, heads=args.heads) * args.steps / (ms_total * 1e-3) / 1e12,
                "train_step": train_step}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
