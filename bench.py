#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): random-walk steps/sec and SGNS word-pairs/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload tract24|tract8|ca|synth100k]

One "step" = one pass of the hot path over one batch of synthetic input of the named shape:
    stage 1  walks over the time-sliced flow graph (+ the spatial walks of the `usespatial` run)
    stage 2  one skip-gram epoch over that walk corpus
Default workload (N=1) = BASELINE.json configs[1]: the Chicago census-tract flow graph, 801 tracts x 24 hourly
layers, 15,000,000 flow walks + 600,000 spatial walks (DeepWalk.java:89-110 sizes), D=20, window=24, K=5.

Prints ONE JSON line.  Top-level metric = walk steps/sec (the metric the reference publishes numbers for); the
line's "stages" object carries both stages, each with value (device-resident), e2e (through the C ABI with host
buffers), roofline and cpu_baseline.  Multi-GPU: one process per GPU (torchrun), walk ids sharded by rank with no
collective; for the tract/CA workloads stage 2 runs as independent replicas ("replicas only").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WALK_BYTES_PER_STEP = 28.0      # row_ptr pair 8 + prob 8 + alias 4 + col 4 + token 4


def sgns_bytes_per_pair(dim, negative):
    return 8.0 * dim * (negative + 2)   # read+write syn0[ctx] and K+1 syn1neg rows, fp32


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at this workload's launch size, from the
    committed ncu capture (profiles/traffic.json: {workload: {kernel: {"bytes": ..., "source": ...}}}), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload, {}).get(kernel, {}).get("bytes")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------ workloads

def make_workload(name, rank=0):
    """Host-side inputs of the named shape (what the Java host would hand to the C ABI)."""
    from embedding_b200 import host, synth
    w = dict(name=name)
    if name in ("tract24", "tract8", "ca"):
        if name == "ca":
            ids, z, L, dim = synth.ca_ids(), synth.ca_latents(), 24, 8
            F = synth.planted_flow_tensor(z, mean_trips_per_pair_hour=2.0)
            n_flow, n_spatial = 8_000_000, 80_000              # DeepWalk.java:93-94,106-107
        else:
            ids, z = synth.tract_ids(), synth.poi_latents()
            L, dim = (24, 20) if name == "tract24" else (8, 20)
            F = synth.planted_flow_tensor(z)
            n_flow, n_spatial = 15_000_000, 600_000            # DeepWalk.java:89-91,102-104
        fl = host.Flows(ids, F)
        host.CrossTimeGraph.numLayer = L
        g = host.CrossTimeGraph.constructGraph_CA(fl) if name == "ca" else host.CrossTimeGraph.constructGraph_tract(fl)
        nv, src, dst, wt = g._bulk
        sp = host.SpatialGraph.constructGraph(ids, synth.planted_spatial_weights(z))
        snv, ssrc, sdst, swt = sp._bulk
        n = len(ids)
        pos = {int(r): i for i, r in enumerate(ids)}
        w.update(L=L, dim=dim, window=L, negative=5, n_regions=n, region_ids=ids,
                 flow=dict(nv=nv, src=src, dst=dst, w=wt, sources=np.array(g.sourceVertices, np.int32),
                           n_walks=n_flow, v_layer=g.v_layer, v_region=g.v_region,
                           id_map=(g.v_layer.astype(np.int64) * n + np.array([pos[int(r)] for r in g.v_region])).astype(np.int32)),
                 spatial=dict(nv=snv, src=ssrc, dst=sdst, w=swt, sources=np.array(sp.sourceVertices, np.int32),
                              out_degree=sp._out_degree_override, sws=sp._sws_override, n_walks=n_spatial,
                              id_map=np.array([pos[int(r)] for r in sp.v_region], np.int32)),
                 n_ids=L * n,
                 desc="%s: %d regions x %d layers, %d flow walks + %d spatial walks, D=%d window=%d K=5"
                      % (name, n, L, n_flow, n_spatial, dim, L))
    elif name == "synth100k":
        n_regions, L = 100_000, 24
        gsyn = synth.powerlaw_flow_graph(n_regions, L=L, seed=100000)   # replicated on every rank
        w.update(L=L, dim=128, window=10, negative=5, n_regions=n_regions,
                 flow=dict(nv=gsyn["n_vertices"], src=gsyn["src"], dst=gsyn["dst"], w=gsyn["w"], sources=gsyn["sources"],
                           n_walks=4_000_000, v_layer=gsyn["v_layer"], v_region=gsyn["v_region"], id_map=None),
                 spatial=None, n_ids=gsyn["n_vertices"],
                 desc="synth100k: 100K regions x 24 slices, %d edges, 4M walks x 24 per GPU, D=128 window=10 K=5" % len(gsyn["src"]))
    else:
        raise SystemExit("unknown workload %r" % name)
    return w


# ------------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        return dict(sm_mhz=float(np.median(busy)) if busy else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU arm

def cpu_reference(w, steps, warmup, walk_sample, sent_sample, threads):
    """The reference's CPU pipeline restated (oracle/): single-threaded alias walks with a java.util.Random LCG
    (CrossTimeGraph.java:134-140), then skip-gram with `threads` Hogwild workers (DeepWalk.java:75).  Returns
    per-step timings on bounded samples of the workload."""
    from oracle import oracle as O
    f = w["flow"]
    g = O.Graph(f["nv"], f["src"], f["dst"], f["w"], f["sources"], alias_mode=O.ALIAS_FAST)
    L = w["L"]
    walk_t, walk_steps, sg_t, sg_pairs = [], [], [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        tok = g.walk(walk_sample, L, seed=100 + it, rng=O.RNG_JAVA_LCG)
        t1 = time.perf_counter()
        sub = tok[:sent_sample]
        p = O.sgns_params(dim=w["dim"], window=w["window"], negative=w["negative"], min_count=2, threads=threads, seed=it)
        m = O.sgns_train(sub, f["nv"], p)
        t2 = time.perf_counter()
        if it >= warmup:
            walk_t.append(t1 - t0)
            walk_steps.append(int((tok >= 0).sum()))
            sg_t.append(t2 - t1)
            sg_pairs.append(m["pairs"])
    return dict(walk_steps_per_s=sum(walk_steps) / sum(walk_t), pairs_per_s=sum(sg_pairs) / sum(sg_t),
                walk_s=sum(walk_t), sgns_s=sum(sg_t), walk_sample=walk_sample, sent_sample=sent_sample)


def run_reference_arm(args, w):
    nproc = os.cpu_count() or 1
    threads = nproc
    r = cpu_reference(w, args.steps, args.warmup, walk_sample=1_000_000, sent_sample=100_000 if w["L"] >= 24 else 600_000,
                      threads=threads)
    ms = (r["walk_s"] + r["sgns_s"]) / args.steps * 1e3
    sample = ("per step: %d single-thread alias walks x L=%d (java.util.Random LCG), then skip-gram over the first %d "
              "of them with %d Hogwild threads; C restatement of the Java pipeline (no JDK in this image)"
              % (r["walk_sample"], w["L"], r["sent_sample"], threads))
    line = dict(impl="reference", metric="walk_steps_per_sec", value=r["walk_steps_per_s"], unit="steps/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64 (alias tables) / f32 (SGNS)", data="synthetic",
                config=dict(workload=w["desc"]),
                e2e=dict(value=r["walk_steps_per_s"], unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                cpu_baseline=dict(value=r["walk_steps_per_s"], unit="steps/s", cores=1, kind="port", sample=sample),
                stages=dict(walk=dict(value=r["walk_steps_per_s"], unit="steps/s", cores=1,
                                      e2e=dict(value=r["walk_steps_per_s"], unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)),
                            sgns=dict(value=r["pairs_per_s"], unit="pairs/s", cores=threads,
                                      e2e=dict(value=r["pairs_per_s"], unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm

def run_gpu_arm(args, w, rank, world, dist):
    from embedding_b200 import abi
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = abi.Context(local)
    data_parallel = dist is not None and w["name"].startswith("synth")
    if data_parallel:      # stage 2 exchanges embedding deltas over NCCL only where the vocabulary is large
        import torch
        from embedding_b200 import parallel
        parallel.init_comm(ctx, dist, torch.device("cuda", local))
    L, dim, neg = w["L"], w["dim"], w["negative"]
    f, sp = w["flow"], w["spatial"]
    peak, peak_src = measured_peak_gbs()

    def sync_all():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def build_graphs():
        G = abi.Graph(ctx, f["nv"], f["src"], f["dst"], f["w"], f["sources"])
        S = None
        if sp is not None:
            S = abi.Graph(ctx, sp["nv"], sp["src"], sp["dst"], sp["w"], sp["sources"], out_degree=sp["out_degree"],
                          source_weight_sum=sp["sws"])
        return G, S

    # walk ids are sharded by rank (weak scaling: every rank samples the full per-GPU count)
    first_flow = rank * f["n_walks"]
    first_sp = rank * (sp["n_walks"] if sp else 0)

    def walk(G, S, seed):
        c1 = G.walk(f["n_walks"], L, seed, first_walk_id=first_flow)
        ms = ctx.phase_ms("walk")
        c2 = None
        if S is not None:
            c2 = S.walk(sp["n_walks"], L, seed + 1, first_walk_id=first_sp)
            ms += ctx.phase_ms("walk")
        return c1, c2, ms

    def relabel(c1, c2):
        if f["id_map"] is not None:
            c1.relabel(f["id_map"], w["n_ids"], 0)
        if c2 is not None:
            c2.relabel(sp["id_map"], w["n_ids"], w["n_regions"])   # "<j>-<region>" tokens, SpatialGraph.java:105-108
        return [c for c in (c1, c2) if c is not None]

    params = abi.sgns_params(dim=dim, window=w["window"], negative=neg, min_count=2, seed=1)

    # ---------------- device-resident timing (value): inputs already in HBM
    G, S = build_graphs()
    launches0 = ctx.kernel_launches()
    walk_ms, walk_kernel_ms, walk_steps, sg_ms, sg_kernel_ms, sg_pairs = [], [], [], [], [], []
    clocks = ClockSampler(local)
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            sync_all()
            if not args.no_clock_sampler:
                clocks.start()
            launches0 = ctx.kernel_launches()
        # device time of each stage: CUDA events on the ctx stream (the stream every libdge kernel runs on)
        ctx.timer_start()
        c1, c2, kms = walk(G, S, seed=1000 + it)
        ev_walk = ctx.timer_stop()
        ctx.timer_start()
        corpora = relabel(c1, c2)
        m = abi.Model.train(ctx, corpora, params)
        ev_sgns = ctx.timer_stop()
        if it >= args.warmup:
            walk_ms.append(ev_walk)
            walk_kernel_ms.append(kms)
            walk_steps.append(sum(c.count_tokens() for c in corpora))
            sg_ms.append(ev_sgns)
            sg_kernel_ms.append(ctx.phase_ms("sgns"))
            sg_pairs.append(m.pairs)
        for c in corpora:
            c.free()
        m.free()
    sync_all()
    clk = clocks.stop()
    launches = ctx.kernel_launches() - launches0
    t_walk = max_over_ranks(sum(walk_ms) / 1e3)
    t_sgns = max_over_ranks(sum(sg_ms) / 1e3)
    tot_steps = sum_over_ranks(float(sum(walk_steps)))
    tot_pairs = sum_over_ranks(float(sum(sg_pairs)))

    # ---------------- end-to-end timing (e2e): host buffers through the C ABI, copies inside the timed region
    n_tok_flow = f["n_walks"] * L
    n_tok_sp = (sp["n_walks"] * L) if sp else 0
    pin_flow = abi.PinnedArray((f["n_walks"], L), np.int32)
    pin_sp = abi.PinnedArray((sp["n_walks"], L), np.int32) if sp else None
    # --tokens16 (off until verified on a GPU): the walk stage hands 16-bit tokens to the host
    tok16 = bool(args.tokens16) and f["nv"] <= 65535 and (sp is None or sp["nv"] <= 65535)
    pin16_flow = abi.PinnedArray((f["n_walks"], L), np.uint16) if tok16 else None
    pin16_sp = abi.PinnedArray((sp["n_walks"], L), np.uint16) if (tok16 and sp) else None
    e_walk_ms, e_sg_ms, e_pairs = [], [], []
    h2d_walk = len(f["src"]) * 16 + len(f["sources"]) * 4 + ((len(sp["src"]) * 16 + sp["nv"] * 12) if sp else 0)
    d2h_walk = (n_tok_flow + n_tok_sp) * (2 if tok16 else 4)
    h2d_sgns = (n_tok_flow + n_tok_sp) * 4
    G.free()
    if S is not None:
        S.free()
    d2h_sgns = 0
    for it in range(0 if args.no_e2e else args.warmup + args.steps):
        if it == args.warmup:
            sync_all()
        t0 = time.perf_counter()
        G, S = build_graphs()                                 # host COO -> device CSR + alias tables
        c1, c2, _ = walk(G, S, seed=2000 + it)
        if tok16:
            c1.tokens_u16(pin16_flow.array)                   # device -> pinned host, 16-bit tokens
            if c2 is not None:
                c2.tokens_u16(pin16_sp.array)
        else:
            c1.tokens(pin_flow.array)                         # device -> pinned host
            if c2 is not None:
                c2.tokens(pin_sp.array)
        t1 = time.perf_counter()
        if tok16:                                             # untimed: the int32 host corpus stage 2 starts from
            c1.tokens(pin_flow.array)
            if c2 is not None:
                c2.tokens(pin_sp.array)
        c1.free()
        if c2 is not None:
            c2.free()
        # stage 2 from HOST tokens (what a Java host holding the corpus would pass)
        t2 = time.perf_counter()
        h1 = abi.Corpus.from_tokens(ctx, pin_flow.array, f["nv"])
        h2 = abi.Corpus.from_tokens(ctx, pin_sp.array, sp["nv"]) if sp else None
        corpora = relabel(h1, h2)
        m = abi.Model.train(ctx, corpora, params)
        syn0, ids = m.vectors()                               # device -> host
        t3 = time.perf_counter()
        if it >= args.warmup:
            e_walk_ms.append((t1 - t0) * 1e3)
            e_sg_ms.append((t3 - t2) * 1e3)
            e_pairs.append(m.pairs)
        d2h_sgns = syn0.nbytes + ids.nbytes
        for c in corpora:
            c.free()
        m.free()
        G.free()
        if S is not None:
            S.free()
    sync_all()
    per_step_steps = sum(walk_steps) / len(walk_steps)        # same expected count per step
    if args.no_e2e:
        e_walk_ms, e_sg_ms, e_pairs = [float("nan")], [float("nan")], [0]
    e_t_walk = max_over_ranks(sum(e_walk_ms) / 1e3)
    e_t_sgns = max_over_ranks(sum(e_sg_ms) / 1e3)
    e_tot_steps = sum_over_ranks(per_step_steps * args.steps)
    e_tot_pairs = sum_over_ranks(float(sum(e_pairs)))

    # ---------------- CPU baseline beside it (rank 0, N=1 only), bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nproc = os.cpu_count() or 1
        # bounded sample, ~20 s of CPU work: ~5 s of single-thread walks, ~15 s of 8-thread skip-gram
        r = cpu_reference(w, 1, 0, walk_sample=10_000_000 if w["name"] != "synth100k" else 4_000_000,
                          sent_sample=(200_000 if w["dim"] >= 64 else 800_000) if L >= 24 else 5_000_000,
                          threads=min(8, nproc))
        cpu = dict(
            walk=dict(value=r["walk_steps_per_s"], unit="steps/s", cores=1, kind="port",
                      sample="%d single-thread alias walks x L=%d with a java.util.Random LCG (oracle/dge_oracle.c, %.1f s)"
                             % (r["walk_sample"], L, r["walk_s"])),
            sgns=dict(value=r["pairs_per_s"], unit="pairs/s", cores=min(8, nproc), kind="port",
                      sample="skip-gram over %d of those walks, %d Hogwild threads as workers(8) (oracle/sgns_oracle.c, %.1f s)"
                             % (r["sent_sample"], min(8, nproc), r["sgns_s"])))

    if rank != 0:
        return
    # which skip-gram kernel the library picked for this shape (phase "sgns_kernel", sgns.cu pick_variant)
    sg_kernel = {0: "k_sgns_seq", 1: "k_sgns_items", 2: "k_sgns_items_v2", 3: "k_sgns_items_g4",
                 4: "k_sgns_items_tp"}.get(int(ctx.phase_ms("sgns_kernel")), "k_sgns_items_v2")
    wk_ms = float(np.mean(walk_kernel_ms))
    sk_ms = float(np.mean(sg_kernel_ms))
    steps_per_launch = per_step_steps
    pairs_per_launch = float(np.mean(sg_pairs))
    walk_ach = steps_per_launch * WALK_BYTES_PER_STEP / (wk_ms * 1e-3) / 1e9
    sgns_ach = pairs_per_launch * sgns_bytes_per_pair(dim, neg) / (sk_ms * 1e-3) / 1e9
    resident_note = ("graph and embedding tables are L2-resident at this size; HBM is not the binding limit (SURVEY 8(d)); the "
                     "binding limit of the item kernel is the L2's 128-bit reduction throughput: 34.5 G rows/s load+reduce "
                     "for this access shape (scripts/red_microbench.cu, profiles/r1s7_red_microbench.txt) = 5.76 G pairs/s"
                     if w["name"] not in ("synth100k", "ca") else
                     "graph and embedding tables are L2-resident (V = 1 848 rows of 32 bytes): neither HBM nor the L2 reduction rate binds; "
                     "the staleness bound (8 * V / (K + 1) = 2 464 pairs in flight, DESIGN.md 3.3) leaves ~8 warps per SM and the epoch "
                     "is latency-bound: pairs / pairs in flight x ~1 700 cycles per pair step (profiles/r1s16_sgns_ca_tp.json)"
                     if w["name"] == "ca" else
                     "tables are 1.2 GB each, but the walk corpus is skewed: most row traffic hits L2 (ncu: 82 % hit rate), so "
                     "algorithmic bytes/s can exceed the HBM peak; see traffic for the measured DRAM bytes")
    stages = dict(
        walk=dict(value=tot_steps / t_walk, unit="steps/s", ms_per_step=t_walk / args.steps * 1e3, kernel="k_walk_alias",
                  kernel_ms=wk_ms,
                  e2e=dict(value=e_tot_steps / e_t_walk, unit="steps/s", h2d_bytes_per_step=int(h2d_walk), d2h_bytes_per_step=int(d2h_walk),
                           includes="dge_graph_build from host COO + dge_walk + dge_corpus_tokens%s to pinned host" % ("_u16" if tok16 else "")),
                  roofline=dict(bound="hbm", achieved=walk_ach, peak=peak, unit="GB/s", frac=walk_ach / peak,
                                traffic=ncu_traffic(w["name"], "k_walk_alias"), bytes_per_unit=WALK_BYTES_PER_STEP, peak_source=peak_src,
                                note=resident_note),
                  cpu_baseline=cpu["walk"] if cpu else None),
        sgns=dict(value=tot_pairs / t_sgns, unit="pairs/s", ms_per_step=t_sgns / args.steps * 1e3, kernel=sg_kernel,
                  kernel_ms=sk_ms, groups_in_flight=ctx.phase_ms("sgns_groups"), sync_rounds=ctx.phase_ms("sgns_rounds"),
                  sync_ms=ctx.phase_ms("sgns_sync"),
                  e2e=dict(value=e_tot_pairs / e_t_sgns, unit="pairs/s", h2d_bytes_per_step=int(h2d_sgns), d2h_bytes_per_step=int(d2h_sgns),
                           includes="dge_corpus_from_tokens from pinned host + dge_sgns_train + dge_model_vectors to host"),
                  roofline=dict(bound="hbm", achieved=sgns_ach, peak=peak, unit="GB/s", frac=sgns_ach / peak,
                                traffic=ncu_traffic(w["name"], "k_sgns_items"), bytes_per_unit=sgns_bytes_per_pair(dim, neg),
                                peak_source=peak_src, note=resident_note),
                  cpu_baseline=cpu["sgns"] if cpu else None))
    share = sk_ms / (sk_ms + wk_ms)
    line = dict(metric="walk_steps_per_sec", value=stages["walk"]["value"], unit="steps/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=(t_walk + t_sgns) / args.steps * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64 (alias tables, walk draws) / f32 (SGNS)", data="synthetic",
                config=dict(workload=w["desc"], l2="inputs of a step (>= 1.5 GB of tokens) exceed the 126 MB L2; every step uses a new seed",
                            parallelism=("1 GPU" if world == 1 else "walk ids sharded by rank, no collective; SGNS data-parallel, NCCL all-reduce of the embedding deltas (per row: sum / contributing ranks)"
                                         if data_parallel else "walk ids sharded by rank, no collective; SGNS replicas only"),
                            timing="CUDA events on the library stream per stage (dge_timer_start/stop), max over ranks"),
                e2e=stages["walk"]["e2e"],
                # roofline of the DOMINANT kernel of the step (the skip-gram item kernel, `share` of the step's kernel
                # time); the walk kernel's own roofline is stages.walk.roofline
                roofline=dict(stages["sgns"]["roofline"], kernel=sg_kernel, share_of_step_kernel_time=share,
                              units="SGNS pairs; the top-level value counts walk steps, see stages"),
                cpu_baseline=stages["walk"]["cpu_baseline"],
                clocks=clk, gpu_launches=int(launches), stages=stages,
                dominant_kernel=dict(name=sg_kernel, share_of_step_kernel_time=share),
                published=dict(note="reference publishes walk wall times only (python/running_time.py:16-20; other hardware, includes "
                                    "String.join + file write): tract alias 0.28 M walks/s, CA alias 0.514 M walks/s = 12.3 M steps/s"))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="tract24")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: device-resident loop only")
    ap.add_argument("--no-clock-sampler", action="store_true", help="experiment: do not poll nvidia-smi during the device-resident loop")
    ap.add_argument("--tokens16", action="store_true", help="walk e2e downloads 16-bit tokens (dge_corpus_tokens_u16; id space < 65536)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0            # the CPU arm runs on rank 0 alone
        run_reference_arm(args, make_workload(args.workload))
        return 0

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        dist = dist_mod
    w = make_workload(args.workload, rank)
    run_gpu_arm(args, w, rank, world, dist)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
