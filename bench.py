#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): random-walk steps/sec and SGNS word-pairs/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload tract24|tract8|ca|synth100k|running_time]

One "step" = one pass of the WHOLE hot path over one batch of synthetic input of the named shape:
    stage 1  walks over the time-sliced flow graph (+ the spatial walks of the `usespatial` run)
    stage 2  one skip-gram epoch over that walk corpus (handed over in device memory)
Default workload = BASELINE.json configs[1]: the Chicago census-tract flow graph, 801 tracts x 24 hourly layers,
15,000,000 flow walks + 600,000 spatial walks (DeepWalk.java:89-110 sizes), D=20, window=24, K=5.

Prints ONE JSON line.  Top-level `value` = walk steps that went through BOTH stages per second (every walk step is
sampled and then trained on): tokens / (walk time + skip-gram time), device-resident; `e2e` = the same through the C ABI
from the host's edge list to the embedding vectors on the host.  `stages` carries each stage's own rate (steps/s,
pairs/s) with roofline, e2e with host buffers, and the CPU port timed beside it.  At N = 1 the line also measures the
HBM-resident synthetic config (BASELINE configs[2], `stages.synth100k`).

Multi-GPU (torchrun, one process per GPU): walk ids sharded by rank with no collective; the tract / CA
skip-gram runs as independent replicas (the tables are KB-MB: "replicas only").  The line's `data_parallel` object is
the synthetic 100K-region workload trained DATA-PARALLEL across the N GPUs -- embedding deltas exchanged by libdge's
peer-memory kernel over NVLink (or NCCL) -- with the sync time, table health and neighbourhood agreement with a
single-GPU run over the same whole corpus.

`--impl reference`: the reference's CPU pipeline restated in C (oracle/; no JDK in this image), timed on the host:
single-thread alias walks with a java.util.Random LCG (CrossTimeGraph.java:134-140), skip-gram on all host threads
(DeepWalk.java:75 asks for workers(8)), both stages over the same bounded sample of the workload per step.
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
METRIC = "walk_steps_per_sec_through_walks_and_skipgram"
SG_KERNELS = {0: "k_sgns_seq", 1: "k_sgns_items", 2: "k_sgns_items_v2", 3: "k_sgns_items_g4", 4: "k_sgns_items_tp",
              5: "k_sgns_items_v3", 6: "k_sgns_items_v2 (plain stores)", 7: "k_sgns_items_v2 (negative table in shared memory)", 8: "k_sgns_sent", 9: "k_sgns_block", 10: "k_sgns_pipe",
              11: "k_sgns_wave", 12: "k_sgns_duo"}


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
    committed ncu capture (profiles/traffic.json: {workload: {kernel: {"bytes", "source", "commit"}}}), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(workload, {}).get(kernel, None)
    except Exception:
        return None


def l2_reduction_ceiling():
    """What the L2 / LSU delivers for the item kernel's access shape, measured LIVE on this GPU by
    scripts/bin/red_microbench (random rows, 128-bit loads + 128-bit reductions): rows/s, or None."""
    exe = os.path.join(ROOT, "scripts", "bin", "red_microbench")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, "--json"], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception:
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

def cpu_reference(w, steps, warmup, n_walks, threads):
    """The reference's CPU pipeline restated (oracle/): single-threaded alias walks with a java.util.Random LCG
    (CrossTimeGraph.java:134-140), then skip-gram over THOSE walks with `threads` Hogwild workers (DeepWalk.java:75).
    Per-step timings on a bounded sample of the workload (n_walks flow walks; the spatial walks in proportion)."""
    from oracle import oracle as O
    f, sp, L = w["flow"], w["spatial"], w["L"]
    g = O.Graph(f["nv"], f["src"], f["dst"], f["w"], f["sources"], alias_mode=O.ALIAS_FAST)
    s = None
    n_sp = 0
    if sp is not None:
        s = O.Graph(sp["nv"], sp["src"], sp["dst"], sp["w"], sp["sources"], out_degree=sp["out_degree"],
                    source_weight_sum=sp["sws"], alias_mode=O.ALIAS_FAST)
        n_sp = max(1, int(round(n_walks * sp["n_walks"] / f["n_walks"])))
    walk_t, walk_steps, sg_t, sg_pairs = [], [], [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        tok = g.walk(n_walks, L, seed=100 + it, rng=O.RNG_JAVA_LCG)
        ts = s.walk(n_sp, L, seed=200 + it, rng=O.RNG_JAVA_LCG) if s is not None else None
        t1 = time.perf_counter()
        # the token strings of the two corpora share one vocabulary ("<h>-<region>"): relabelled as on the GPU arm (the Java
        # host gets this for free from its String tokens, so it is outside both timed stages)
        if f["id_map"] is not None:
            tok = np.where(tok >= 0, f["id_map"][np.maximum(tok, 0)], -1).astype(np.int32)
        if ts is not None:
            pos = np.arange(L, dtype=np.int64)[None, :] * w["n_regions"]
            ts = np.where(ts >= 0, sp["id_map"][np.maximum(ts, 0)] + pos, -1).astype(np.int32)
            tok = np.concatenate([tok, ts])
        p = O.sgns_params(dim=w["dim"], window=w["window"], negative=w["negative"], min_count=2, threads=threads, seed=it)
        t2 = time.perf_counter()
        m = O.sgns_train(tok, w["n_ids"], p)
        t3 = time.perf_counter()
        if it >= warmup:
            walk_t.append(t1 - t0)
            walk_steps.append(int((tok >= 0).sum()))
            sg_t.append(t3 - t2)
            sg_pairs.append(m["pairs"])
    return dict(walk_steps_per_s=sum(walk_steps) / sum(walk_t), pairs_per_s=sum(sg_pairs) / sum(sg_t),
                whole_steps_per_s=sum(walk_steps) / (sum(walk_t) + sum(sg_t)),
                walk_s=sum(walk_t), sgns_s=sum(sg_t), n_walks=n_walks, n_spatial=n_sp, steps=sum(walk_steps), pairs=sum(sg_pairs))


def cpu_sample_text(r, w, threads):
    return ("per step: %d flow%s walks x L=%d sampled by ONE thread (alias method, java.util.Random LCG, CrossTimeGraph.java:134-140), then one "
            "skip-gram epoch over those walks on %d Hogwild threads (DeepWalk.java:75 asks for workers(8)); C restatement of "
            "the Java pipeline (oracle/, no JDK in this image); a SAMPLE of the workload's %d walks -- rates, not totals, are compared"
            % (r["n_walks"], (" + %d spatial" % r["n_spatial"]) if r["n_spatial"] else "", w["L"], threads, w["flow"]["n_walks"]))


def run_reference_arm(args, w):
    threads = os.cpu_count() or 1      # every host thread the box has (the reference asks for 8 workers)
    n_walks = 150_000 if w["L"] >= 24 and w["dim"] < 64 else (60_000 if w["dim"] >= 64 else 600_000)
    r = cpu_reference(w, args.steps, args.warmup, n_walks, threads)
    ms = (r["walk_s"] + r["sgns_s"]) / args.steps * 1e3
    sample = cpu_sample_text(r, w, threads)
    e2e = dict(value=r["whole_steps_per_s"], unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    line = dict(impl="reference", metric=METRIC, value=r["whole_steps_per_s"], unit="steps/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64 (alias tables, walk draws) / f32 (SGNS)", data="synthetic",
                config=dict(workload=w["desc"], sample=sample), e2e=e2e,
                cpu_baseline=dict(value=r["whole_steps_per_s"], unit="steps/s", cores=threads, kind="port", sample=sample),
                stages=dict(walk=dict(value=r["walk_steps_per_s"], unit="steps/s", cores=1, seconds=r["walk_s"]),
                            sgns=dict(value=r["pairs_per_s"], unit="pairs/s", cores=threads, seconds=r["sgns_s"])),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm

class Dist:
    """barrier / max / sum over ranks through torch.distributed (plumbing only); identity at N = 1."""

    def __init__(self, dist):
        self.dist = dist

    def sync(self):
        if self.dist is not None:
            import torch
            self.dist.barrier()
            torch.cuda.synchronize()

    def _red(self, x, op):
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX) if self.dist is not None else x

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM) if self.dist is not None else x


def build_graphs(abi, ctx, w):
    f, sp = w["flow"], w["spatial"]
    G = abi.Graph(ctx, f["nv"], f["src"], f["dst"], f["w"], f["sources"])
    S = None
    if sp is not None:
        S = abi.Graph(ctx, sp["nv"], sp["src"], sp["dst"], sp["w"], sp["sources"], out_degree=sp["out_degree"],
                      source_weight_sum=sp["sws"])
    return G, S


def walk_both(ctx, w, G, S, seed, rank):
    f, sp, L = w["flow"], w["spatial"], w["L"]
    c1 = G.walk(f["n_walks"], L, seed, first_walk_id=rank * f["n_walks"])       # walk ids sharded by rank (weak scaling)
    ms = ctx.phase_ms("walk")
    c2 = None
    if S is not None:
        c2 = S.walk(sp["n_walks"], L, seed + 1, first_walk_id=rank * sp["n_walks"])
        ms += ctx.phase_ms("walk")
    return c1, c2, ms


def relabel(w, c1, c2):
    f, sp = w["flow"], w["spatial"]
    if f["id_map"] is not None:
        c1.relabel(f["id_map"], w["n_ids"], 0)
    if c2 is not None:
        c2.relabel(sp["id_map"], w["n_ids"], w["n_regions"])   # "<j>-<region>" tokens, SpatialGraph.java:105-108
    return [c for c in (c1, c2) if c is not None]


def run_main_workload(args, w, rank, world, D, abi, ctx, local):
    """The whole hot path on workload `w`: device-resident loop (value), end-to-end loop (e2e), stage e2e loops."""
    L, dim, neg = w["L"], w["dim"], w["negative"]
    f, sp = w["flow"], w["spatial"]
    params = abi.sgns_params(dim=dim, window=w["window"], negative=neg, min_count=2, seed=1)
    G, S = build_graphs(abi, ctx, w)
    walk_ms, walk_kernel_ms, tokens, sg_ms, sg_kernel_ms, sg_pairs = [], [], [], [], [], []
    clocks = ClockSampler(local)
    launches0 = ctx.kernel_launches()
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            D.sync()
            if not args.no_clock_sampler:
                clocks.start()
            launches0 = ctx.kernel_launches()
        # device time of each stage: CUDA events on the ctx stream (the stream every libdge kernel runs on)
        ctx.timer_start()
        c1, c2, kms = walk_both(ctx, w, G, S, 1000 + it, rank)
        ev_walk = ctx.timer_stop()
        ctx.timer_start()
        corpora = relabel(w, c1, c2)
        m = abi.Model.train(ctx, corpora, params)
        ev_sgns = ctx.timer_stop()
        if it >= args.warmup:
            walk_ms.append(ev_walk)
            walk_kernel_ms.append(kms)
            tokens.append(sum(c.count_tokens() for c in corpora))
            sg_ms.append(ev_sgns)
            sg_kernel_ms.append(ctx.phase_ms("sgns"))
            sg_pairs.append(m.pairs)
        for c in corpora:
            c.free()
        m.free()
    D.sync()
    clk = clocks.stop() if not args.no_clock_sampler else dict(sm_mhz=None, sm_max_mhz=None, reasons=["sampler disabled"])
    launches = ctx.kernel_launches() - launches0
    sg_kernel = SG_KERNELS.get(int(ctx.phase_ms("sgns_kernel")), "k_sgns_items_v2")
    groups = ctx.phase_ms("sgns_groups")
    write_through = int(ctx.phase_ms("sgns_write_through"))
    t_walk = D.max(sum(walk_ms) / 1e3)
    t_sgns = D.max(sum(sg_ms) / 1e3)
    t_step = D.max((sum(walk_ms) + sum(sg_ms)) / 1e3)
    tot_tokens = D.sum(float(sum(tokens)))
    tot_pairs = D.sum(float(sum(sg_pairs)))
    per_step_tokens = sum(tokens) / len(tokens)
    G.free()
    if S is not None:
        S.free()

    # ---------------- end to end: host edge lists -> embedding vectors on the host, through the C ABI
    coo_bytes = len(f["src"]) * 16 + len(f["sources"]) * 4 + ((len(sp["src"]) * 16 + sp["nv"] * 12) if sp else 0)
    e_ms, vec_bytes = [], 0
    for it in range(0 if args.no_e2e else args.warmup + args.steps):
        if it == args.warmup:
            D.sync()
        t0 = time.perf_counter()
        G, S = build_graphs(abi, ctx, w)                      # host COO -> device CSR + alias tables + walk records
        c1, c2, _ = walk_both(ctx, w, G, S, 2000 + it, rank)
        corpora = relabel(w, c1, c2)                          # the corpus is handed to stage 2 in device memory
        m = abi.Model.train(ctx, corpora, params)
        syn0, ids = m.vectors()                               # device -> host: what writeWordVectors serialises
        t1 = time.perf_counter()
        if it >= args.warmup:
            e_ms.append((t1 - t0) * 1e3)
        vec_bytes = syn0.nbytes + ids.nbytes
        for c in corpora:
            c.free()
        m.free()
        G.free()
        if S is not None:
            S.free()
    D.sync()
    e2e = None
    if not args.no_e2e:
        e_t = D.max(sum(e_ms) / 1e3)
        e2e = dict(value=D.sum(per_step_tokens * args.steps) / e_t, unit="steps/s", h2d_bytes_per_step=int(coo_bytes),
                   d2h_bytes_per_step=int(vec_bytes), ms_per_step=e_t / args.steps * 1e3,
                   includes="both stages: dge_graph_build from the host edge lists (flow + spatial graph), dge_walk, dge_corpus_relabel, "
                            "dge_sgns_train on the device corpus, dge_model_vectors to the host; host wall clock around the blocking calls, max over ranks")

    # ---------------- per-stage end to end with HOST buffers on both sides of each stage (a host that wants the tokens)
    st = min(3, args.steps)
    tok16 = (not args.tokens32) and f["nv"] <= 65535 and (sp is None or sp["nv"] <= 65535)
    stage_e2e = None
    if not args.no_e2e:
        dt = np.uint16 if tok16 else np.int32
        pin_flow = abi.PinnedArray((f["n_walks"], L), dt)
        pin_sp = abi.PinnedArray((sp["n_walks"], L), dt) if sp else None
        pin32_flow = abi.PinnedArray((f["n_walks"], L), np.int32) if tok16 else pin_flow
        pin32_sp = (abi.PinnedArray((sp["n_walks"], L), np.int32) if sp else None) if tok16 else pin_sp
        ew, es, ep = [], [], []
        for it in range(1 + st):
            t0 = time.perf_counter()
            G, S = build_graphs(abi, ctx, w)
            c1, c2, _ = walk_both(ctx, w, G, S, 3000 + it, rank)
            if tok16:                                         # 16-bit tokens for id spaces below 65 535: half the PCIe bytes
                c1.tokens_u16(pin_flow.array)
                if c2 is not None:
                    c2.tokens_u16(pin_sp.array)
            else:
                c1.tokens(pin_flow.array)
                if c2 is not None:
                    c2.tokens(pin_sp.array)
            t1 = time.perf_counter()
            if tok16:                                         # untimed: the int32 host corpus stage 2 starts from
                c1.tokens(pin32_flow.array)
                if c2 is not None:
                    c2.tokens(pin32_sp.array)
            c1.free()
            if c2 is not None:
                c2.free()
            t2 = time.perf_counter()
            h1 = abi.Corpus.from_tokens(ctx, pin32_flow.array, f["nv"])
            h2 = abi.Corpus.from_tokens(ctx, pin32_sp.array, sp["nv"]) if sp else None
            corpora = relabel(w, h1, h2)
            m = abi.Model.train(ctx, corpora, params)
            m.vectors()
            t3 = time.perf_counter()
            if it >= 1:
                ew.append(t1 - t0)
                es.append(t3 - t2)
                ep.append(m.pairs)
            for c in corpora:
                c.free()
            m.free()
            G.free()
            if S is not None:
                S.free()
        n_tok_all = (f["n_walks"] + (sp["n_walks"] if sp else 0)) * L
        stage_e2e = dict(
            walk=dict(value=D.sum(per_step_tokens * st) / D.max(sum(ew)), unit="steps/s", h2d_bytes_per_step=int(coo_bytes),
                      d2h_bytes_per_step=int(n_tok_all * (2 if tok16 else 4)), steps=st,
                      includes="dge_graph_build from host COO + dge_walk + dge_corpus_tokens%s to pinned host" % ("_u16" if tok16 else "")),
            sgns=dict(value=D.sum(float(sum(ep))) / D.max(sum(es)), unit="pairs/s", h2d_bytes_per_step=int(n_tok_all * 4),
                      d2h_bytes_per_step=int(vec_bytes), steps=st,
                      includes="dge_corpus_from_tokens from pinned host + dge_sgns_train + dge_model_vectors to host"))
    return dict(t_walk=t_walk, t_sgns=t_sgns, t_step=t_step, tot_tokens=tot_tokens, tot_pairs=tot_pairs,
                per_step_tokens=per_step_tokens, pairs_per_launch=float(np.mean(sg_pairs)),
                wk_ms=float(np.mean(walk_kernel_ms)), sk_ms=float(np.mean(sg_kernel_ms)), sg_kernel=sg_kernel, groups=groups, write_through=write_through,
                clocks=clk, launches=launches, e2e=e2e, stage_e2e=stage_e2e)


def knn_overlap_sample(a, b, n_query=2000, n_cand=50_000, k=10):
    """Mean overlap of the k nearest cosine neighbours of `n_query` rows among the first `n_cand` rows (the most frequent words)."""
    n = min(n_cand, len(a), len(b))
    A, B = a[:n].astype(np.float32), b[:n].astype(np.float32)
    A /= np.maximum(np.linalg.norm(A, axis=1, keepdims=True), 1e-12)
    B /= np.maximum(np.linalg.norm(B, axis=1, keepdims=True), 1e-12)
    q = np.linspace(0, n - 1, min(n_query, n)).astype(np.int64)
    hits = 0
    for lo in range(0, len(q), 250):
        qq = q[lo:lo + 250]
        sa, sb = A[qq] @ A.T, B[qq] @ B.T
        sa[np.arange(len(qq)), qq] = -np.inf
        sb[np.arange(len(qq)), qq] = -np.inf
        na = np.argpartition(-sa, k, axis=1)[:, :k]
        nb = np.argpartition(-sb, k, axis=1)[:, :k]
        hits += sum(len(np.intersect1d(x, y)) for x, y in zip(na, nb))
    return round(hits / (len(q) * k), 4)


_CEIL = {}


def synth_sgns_roofline(ach_gbs, pairs_per_s, dim, neg, peak, peak_src, traffic, kernel_ms):
    """The D = 128 tables are 1.2 GB each, but the walk corpus is skewed and most row traffic hits the 126 MB L2 (ncu: DRAM
    traffic ~ 5 % of the algorithmic bytes), so HBM does not bound the kernel: its limit is what the L2 delivers for
    512-byte-row 128-bit loads + reductions on the hot set, measured live by scripts/bin/red_microbench."""
    if "c" not in _CEIL:
        _CEIL["c"] = l2_reduction_ceiling()
    c = _CEIL["c"]
    rows = c["d128_hot100k_rows_per_s"]["load_red"] if c and "d128_hot100k_rows_per_s" in c else None
    bpp = sgns_bytes_per_pair(dim, neg)
    ceil_pairs = rows / (neg + 1) if rows else None
    return dict(bound="l2-reduction", achieved=ach_gbs, peak=(ceil_pairs * bpp / 1e9) if ceil_pairs else None, unit="GB/s",
                frac=(pairs_per_s / ceil_pairs) if ceil_pairs else None, bytes_per_unit=bpp,
                peak_source=("scripts/bin/red_microbench --json, live: %.3g rows/s of 512 bytes loaded AND reduced on a 100 000-row hot set / (K + 1) rows per "
                             "pair = %.3g pairs/s; the kernel can exceed it because it sends no reduction for a saturated sigmoid (g = 0)" % (rows, ceil_pairs)) if rows else "micro-benchmark binary missing",
                traffic=traffic["bytes"] if traffic else None,
                hbm=dict(peak=peak, peak_source=peak_src, frac_by_algorithmic_bytes=ach_gbs / peak,
                         frac_by_dram_traffic=(traffic["bytes"] / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         note="algorithmic bytes/s exceed the HBM peak because ~80-95 % of the row traffic hits L2; the DRAM-traffic fraction is the HBM load"))


def run_synth(args, rank, world, D, dist, abi, ctx, local, steps=3, warmup=1):
    """BASELINE configs[2] (HBM-resident): 100K regions x 24 slices, ~100M edges, 4M walks x 24 per GPU, D=128.
    N = 1: one GPU.  N > 1: walk ids sharded by rank, skip-gram DATA-PARALLEL over the ranks' corpus shards."""
    t_gen = time.time()
    w = make_workload("synth100k", rank)
    gen_s = time.time() - t_gen
    f, L, dim, neg = w["flow"], w["L"], w["dim"], w["negative"]
    peak, peak_src = measured_peak_gbs()
    multi = dist is not None and world > 1
    if multi:
        import torch
        from embedding_b200 import parallel
        parallel.init_comm(ctx, dist, torch.device("cuda", local))
    params = abi.sgns_params(dim=dim, window=w["window"], negative=neg, min_count=2, seed=1, sync_rounds=args.sync_rounds,
                             transport=args.transport, combine=args.combine)
    t0 = time.perf_counter()
    G = abi.Graph(ctx, f["nv"], f["src"], f["dst"], f["w"], f["sources"])
    build_s = time.perf_counter() - t0
    wms, sms, toks, pairs, syncs, phases = [], [], [], [], [], []
    stats = rounds = transport = keep = None
    for it in range(warmup + steps):
        if it == warmup:
            D.sync()
        last = it == warmup + steps - 1
        ctx.timer_start()
        c = G.walk(f["n_walks"], L, 777 if last else 1000 + it, first_walk_id=rank * f["n_walks"])
        ev_w = ctx.timer_stop()
        kms = ctx.phase_ms("walk")
        ctx.timer_start()
        m = abi.Model.train(ctx, [c], params)
        ev_s = ctx.timer_stop()
        if it >= warmup:
            wms.append((ev_w, kms))
            sms.append((ev_s, ctx.phase_ms("sgns")))
            toks.append(c.count_tokens())
            pairs.append(m.pairs)
            syncs.append(ctx.phase_ms("sgns_sync"))
            phases.append(dict(vocab_ms=ctx.phase_ms("vocab"), compact_ms=ctx.phase_ms("compact"), train_ms=ctx.phase_ms("sgns"),
                               exchange_setup_ms=ctx.phase_ms("sgns_dp_setup")))
        rounds = ctx.phase_ms("sgns_rounds")
        transport = {0.0: "none", 1.0: "peer-memory kernel over NVLink (cudaIpc)", 2.0: "NCCL all-reduce"}.get(ctx.phase_ms("sgns_transport"), "?")
        if last:
            stats = m.stats()
            keep = m                                             # the model of walk seed 777 (compared with one GPU below)
        else:
            m.free()
        c.free()
    D.sync()
    sg_kernel = SG_KERNELS.get(int(ctx.phase_ms("sgns_kernel")), "?")
    t_walk, t_sgns = D.max(sum(x[0] for x in wms) / 1e3), D.max(sum(x[0] for x in sms) / 1e3)
    tot_tok, tot_pairs = D.sum(float(sum(toks))), D.sum(float(sum(pairs)))
    wk_ms, sk_ms = float(np.mean([x[1] for x in wms])), float(np.mean([x[1] for x in sms]))
    walk_ach = float(np.mean(toks)) * WALK_BYTES_PER_STEP / (wk_ms * 1e-3) / 1e9
    sg_ach = float(np.mean(pairs)) * sgns_bytes_per_pair(dim, neg) / (sk_ms * 1e-3) / 1e9
    tw, ts = ncu_traffic("synth100k", "k_walk_alias"), ncu_traffic("synth100k", "sgns")
    out = dict(
        workload=w["desc"], host_generation_s=round(gen_s, 1), graph_build_from_host_coo_s=round(build_s, 2), steps=steps, warmup=warmup,
        value=tot_tok / (t_walk + t_sgns), unit="steps/s", ms_per_step=(t_walk + t_sgns) / steps * 1e3,
        walk=dict(value=tot_tok / t_walk, unit="steps/s", kernel="k_walk_alias", kernel_ms=wk_ms,
                  roofline=dict(bound="hbm", achieved=walk_ach, peak=peak, unit="GB/s", frac=walk_ach / peak, bytes_per_unit=WALK_BYTES_PER_STEP,
                                peak_source=peak_src, traffic=tw["bytes"] if tw else None,
                                note="3.2 GB of 32-byte walk records, 25x the L2: one dependent random sector per step; the hardware's own rate "
                                     "for this access shape is measured by scripts/walk_microbench.cu (profiles/)")),
        sgns=dict(value=tot_pairs / t_sgns, unit="pairs/s", kernel=sg_kernel, kernel_ms=sk_ms, sync_rounds=rounds, sync_ms=float(np.mean(syncs)),
                  transport=transport, call_ms=t_sgns / steps * 1e3,
                  call_phases_ms={k: float(np.mean([ph[k] for ph in phases])) for k in phases[0]},
                  kernel_pairs_per_s=tot_pairs / D.max(sum(x[1] for x in sms) / 1e3),
                  roofline=synth_sgns_roofline(sg_ach, float(np.mean(pairs)) / (sk_ms * 1e-3), dim, neg, peak, peak_src, ts, sk_ms)),
        model_stats=stats)
    if multi:
        # ---- what ONE GPU does with the same per-GPU work, in the same run (no communicator): the scaling reference
        single = overlap = None
        ctx1 = abi.Context(local)
        if rank == 0:
            G1 = abi.Graph(ctx1, f["nv"], f["src"], f["dst"], f["w"], f["sources"])
            p1 = abi.sgns_params(dim=dim, window=w["window"], negative=neg, min_count=2, seed=1)
            ms1, pr1, cl1 = [], [], []
            for it in range(3):
                c = G1.walk(f["n_walks"], L, 1000 + it, first_walk_id=0)
                ctx1.timer_start()
                m1 = abi.Model.train(ctx1, [c], p1)
                call = ctx1.timer_stop()
                if it >= 1:
                    ms1.append(ctx1.phase_ms("sgns"))
                    cl1.append(call)
                    pr1.append(m1.pairs)
                m1.free()
                c.free()
            single = dict(value=sum(pr1) / (sum(cl1) / 1e3), unit="pairs/s", kernel_pairs_per_s=sum(pr1) / (sum(ms1) / 1e3), call_ms=float(np.mean(cl1)),
                          note="one GPU of this box, the same 4M walks per GPU, no communicator, same run: `value` times the whole dge_sgns_train "
                               "call like sgns.value above; kernel_pairs_per_s the training launches only")
            # ---- neighbourhood agreement of the data-parallel embedding with a single-GPU run over the WHOLE corpus (walk
            # ids [0, N x 4M) of seed 777: exactly the union of the ranks' shards), and the single-GPU noise floor
            cu = G1.walk(f["n_walks"] * world, L, 777, first_walk_id=0)
            mu = abi.Model.train(ctx1, [cu], p1)
            mu2 = abi.Model.train(ctx1, [cu], abi.sgns_params(dim=dim, window=w["window"], negative=neg, min_count=2, seed=2))
            s_dp, ids_dp = keep.vectors()
            s_u, ids_u = mu.vectors()
            s_u2, _ = mu2.vectors()
            overlap = dict(knn_overlap_vs_single_gpu=knn_overlap_sample(s_u, s_dp), knn_overlap_two_single_gpu_seeds=knn_overlap_sample(s_u, s_u2),
                           same_vocabulary=bool(np.array_equal(ids_dp, ids_u)), pairs_single_gpu=int(mu.pairs),
                           note="k = 10 cosine neighbours of 2000 words among the 50 000 most frequent; single-GPU run over the union of the "
                                "ranks' walk shards, same seed; second number = two single-GPU runs with different seeds (noise floor)")
            for x in (mu, mu2, cu, G1):
                x.free()
        dp_pairs = int(D.sum(float(keep.pairs)))
        if overlap is not None:
            overlap["pairs_data_parallel"] = dp_pairs
        D.sync()
        out["single_gpu_reference"] = single
        out["agreement"] = overlap
        ctx1.close()
    keep.free()
    G.free()
    return out


def run_gpu_arm(args, w, rank, world, dist):
    from embedding_b200 import abi
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = abi.Context(local)
    D = Dist(dist)
    L, dim, neg = w["L"], w["dim"], w["negative"]
    peak, peak_src = measured_peak_gbs()
    if args.dp_only:
        out = run_synth(args, rank, world, D, dist, abi, ctx, local)
        if rank == 0:
            print(json.dumps(dict(n_gpus=world, sync_rounds_arg=args.sync_rounds, transport_arg=args.transport, combine_arg=args.combine, **out)), flush=True)
        return
    R = run_main_workload(args, w, rank, world, D, abi, ctx, local)

    # ---------------- CPU baseline beside it (rank 0, N=1 only), bounded sample (~20 s)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = min(8, os.cpu_count() or 1)                  # workers(8), DeepWalk.java:75
        r = cpu_reference(w, 1, 0, 400_000 if (L >= 24 and dim < 64) else (100_000 if dim >= 64 else 3_000_000), threads)
        txt = cpu_sample_text(r, w, threads)
        cpu = dict(whole=dict(value=r["whole_steps_per_s"], unit="steps/s", cores=threads, kind="port", sample=txt + " (%.1f s)" % (r["walk_s"] + r["sgns_s"])),
                   walk=dict(value=r["walk_steps_per_s"], unit="steps/s", cores=1, kind="port", sample="the walk stage of that sample (%.1f s)" % r["walk_s"]),
                   sgns=dict(value=r["pairs_per_s"], unit="pairs/s", cores=threads, kind="port", sample="the skip-gram stage of that sample (%.1f s)" % r["sgns_s"]))

    # ---------------- the HBM-resident synthetic config: one GPU at N = 1, data-parallel at N > 1
    synth = None
    if not args.no_synth and w["name"] != "synth100k":
        synth = run_synth(args, rank, world, D, dist, abi, ctx, local)
    if rank != 0:
        return

    ceil = l2_reduction_ceiling() if w["name"] != "synth100k" else None
    walk_ach = R["per_step_tokens"] * WALK_BYTES_PER_STEP / (R["wk_ms"] * 1e-3) / 1e9
    bpp = sgns_bytes_per_pair(dim, neg)
    sgns_ach = R["pairs_per_launch"] * bpp / (R["sk_ms"] * 1e-3) / 1e9
    sg_traffic = ncu_traffic(w["name"], "sgns")
    if w["name"] == "synth100k":
        sg_roof = dict(bound="hbm", achieved=sgns_ach, peak=peak, unit="GB/s", frac=sgns_ach / peak, peak_source=peak_src)
    else:
        # L2-resident tables: HBM is not the limit (ncu: DRAM traffic is a fraction of a percent of the algorithmic bytes); the
        # limit is what the LSU / L2 delivers for random-row 128-bit loads + 128-bit reductions, measured live
        key = "ca_rows_per_s" if w["name"] == "ca" else "tract24_rows_per_s"
        rows = ceil[key]["load_red"] if ceil else (34.5e9 if w["name"] != "ca" else None)
        ceil_pairs = rows / (neg + 1) if rows else None
        sg_roof = dict(bound="l2-reduction", achieved=sgns_ach, peak=(ceil_pairs * bpp / 1e9) if ceil_pairs else None, unit="GB/s",
                       frac=(sgns_ach / (ceil_pairs * bpp / 1e9)) if ceil_pairs else None,
                       peak_source=("scripts/bin/red_microbench --json, run live on this GPU: %.3g rows/s loaded AND reduced (128-bit lanes, this row shape) "
                                    "/ (K + 1) rows per pair = %.3g pairs/s" % (rows, ceil_pairs)) if ceil else "profiles/r1s7_red_microbench.txt (micro-benchmark binary missing)",
                       hbm=dict(peak=peak, peak_source=peak_src, frac_by_algorithmic_bytes=sgns_ach / peak,
                                frac_by_dram_traffic=(sg_traffic["bytes"] / (R["sk_ms"] * 1e-3) / 1e9 / peak) if sg_traffic else None,
                                note="tables (%.1f MB) and corpus windows are L2-resident: the HBM figure is not the binding limit" % (2 * w["n_ids"] * 4 * ((dim + 7) // 8 * 8) / 1e6)))
    sg_roof.update(traffic=sg_traffic["bytes"] if sg_traffic else None, traffic_source=sg_traffic.get("source") if sg_traffic else None,
                   bytes_per_unit=bpp, kernel=R["sg_kernel"])
    wk_traffic = ncu_traffic(w["name"], "k_walk_alias")
    se = R["stage_e2e"] or {}
    stages = dict(
        walk=dict(value=R["tot_tokens"] / R["t_walk"], unit="steps/s", ms_per_step=R["t_walk"] / args.steps * 1e3, kernel="k_walk_alias",
                  kernel_ms=R["wk_ms"], e2e=se.get("walk"),
                  roofline=dict(bound="hbm" if w["name"] == "synth100k" else "l2", achieved=walk_ach, peak=peak, unit="GB/s", frac=walk_ach / peak,
                                traffic=wk_traffic["bytes"] if wk_traffic else None, bytes_per_unit=WALK_BYTES_PER_STEP, peak_source=peak_src,
                                note="" if w["name"] == "synth100k" else "the record array is L2-resident at this size; the HBM peak is quoted as the contract's "
                                     "denominator, the kernel is bound by L1/L2 request throughput of one 32-byte sector per step"),
                  cpu_baseline=cpu["walk"] if cpu else None),
        sgns=dict(value=R["tot_pairs"] / R["t_sgns"], unit="pairs/s", ms_per_step=R["t_sgns"] / args.steps * 1e3, kernel=R["sg_kernel"],
                  kernel_ms=R["sk_ms"], groups_in_flight=R["groups"], write_through_words=R["write_through"],
                  schedule=("a warp per sentence, sentences handed out from a counter, the %d most frequent words written through" % R["write_through"]) if R["sg_kernel"] == "k_sgns_sent" else None,
                  e2e=se.get("sgns"), roofline=sg_roof,
                  cpu_baseline=cpu["sgns"] if cpu else None))
    if synth is not None and world == 1:
        stages["synth100k"] = synth
    share = R["sk_ms"] / (R["sk_ms"] + R["wk_ms"])
    line = dict(metric=METRIC, value=R["tot_tokens"] / R["t_step"], unit="steps/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=R["t_step"] / args.steps * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64 (alias tables, walk draws) / f32 (SGNS)", data="synthetic",
                config=dict(workload=w["desc"],
                            counts="value = walk steps sampled AND trained on per second: tokens / (walk + skip-gram device time); stages has steps/s and pairs/s",
                            l2="inputs of a step (>= 1.5 GB of tokens) exceed the 126 MB L2; every step uses a new seed",
                            parallelism=("1 GPU" if world == 1 else "walk ids sharded by rank, no collective; skip-gram replicas only on this workload (KB-MB tables); "
                                         "the data-parallel skip-gram over NVLink is measured on the synthetic workload: see data_parallel"),
                            timing="CUDA events on the library stream per stage (dge_timer_start/stop), max over ranks"),
                e2e=R["e2e"], roofline=dict(sg_roof, share_of_step_kernel_time=share, units="SGNS pairs (the step's dominant kernel)"),
                cpu_baseline=cpu["whole"] if cpu else None, clocks=R["clocks"], gpu_launches=int(R["launches"]), stages=stages,
                dominant_kernel=dict(name=R["sg_kernel"], share_of_step_kernel_time=share),
                published=dict(note="the reference publishes walk wall times only (python/running_time.py:16-20; other hardware, includes String.join + "
                                    "file write): tract alias 0.28 M walks/s, CA alias 0.514 M walks/s = 12.3 M steps/s; `bench.py --workload running_time` reproduces that table"))
    if synth is not None and world > 1:
        line["data_parallel"] = synth
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
    ap.add_argument("--no-synth", action="store_true", help="skip the synthetic 100K-region stage pair / data-parallel run")
    ap.add_argument("--dp-only", action="store_true", help="sweeps: only the synthetic stage pair / data-parallel run; prints that object")
    ap.add_argument("--no-clock-sampler", action="store_true", help="experiment: do not poll nvidia-smi during the device-resident loop")
    ap.add_argument("--tokens32", action="store_true", help="stage e2e of the walks downloads int32 tokens even when the id space fits 16 bits")
    ap.add_argument("--sync-rounds", type=int, default=0, help="data-parallel skip-gram: exchanges per epoch (0 = automatic)")
    ap.add_argument("--transport", type=int, default=0, help="0 auto, 1 peer-memory kernel, 2 NCCL")
    ap.add_argument("--combine", type=int, default=0, help="dge.h DGE_COMBINE_* (0 = default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.workload == "running_time":
        from scripts import running_time
        return running_time.main(args)
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
    w = make_workload(args.workload, rank) if not args.dp_only else dict(name="none", L=0, dim=0, negative=0)
    run_gpu_arm(args, w, rank, world, dist)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
