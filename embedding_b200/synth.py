"""Synthetic inputs of the named shapes (BASELINE.json configs).

The reference ships no flow data, so every config runs on seeded synthetic flows:
  * `flow_tensor(n, seed, density)`            F[src, hour(24), dst] int32 trip counts (CA 77 / tract 801)
  * `powerlaw_flow_graph(n_regions, L, seed)`  COO of a time-sliced graph with Zipf out-degrees and
                                               Pareto integer weights (100K / 1M regions x 24 slices)
  * `spatial_weights(n, seed)`                 exp(-100 d) over random centroids (SpatialGraph.java:43-48)
numpy only; used by tests, bench.py and the host mirror.
"""
import json
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def tract_ids():
    """The 801 real Chicago tract ids (fixture extracted from the reference's miscs/POI_tract.pickle)."""
    with open(os.path.join(_GOLDEN, "poi_tract.json")) as f:
        return np.array(json.load(f)["tract_ids"], np.int32)


def ca_ids():
    return np.arange(1, 78, dtype=np.int32)  # community areas are numbered 1..77 (CommunityAreas.java:210)


def flow_tensor(n, seed=2013, density=0.6, alpha=1.2, scale=3.0):
    """F[src, hour, dst] int32: floor(Pareto(alpha) * scale) masked to `density`, heavier in daytime hours."""
    rng = np.random.default_rng(seed)
    F = np.floor(rng.pareto(alpha, size=(n, 24, n)) * scale).astype(np.int64)
    mask = rng.random((n, 24, n)) < density
    hour_profile = 0.35 + 0.65 * np.sin(np.linspace(0, np.pi, 24)) ** 2
    F = np.floor(F * mask * hour_profile[None, :, None]).astype(np.int64)
    return np.minimum(F, 2 ** 20).astype(np.int32)


def spatial_weights(n, seed=2013, extent=0.25):
    """w = exp(-100 * centroid distance) over `n` random 2-D centroids (degrees, like the shapefile's)."""
    rng = np.random.default_rng(seed)
    xy = rng.random((n, 2)) * extent
    d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    return np.exp(-d * 100.0)


def planted_spatial_weights(z, seed=2013, extent=0.25, jitter=0.02):
    """w = exp(-100 * centroid distance) (SpatialGraph.java:43-48) over centroids PLANTED from the regions' latent
    vectors z: the two leading principal directions of the normalised latents, scaled into `extent` degrees, plus a
    little jitter -- regions with similar POI / label profiles are neighbours, as in a real city, so the spatial
    corpus of the `usespatial` run carries the same ground truth as the planted flows (with independent random
    centroids its 10-nearest-neighbour walks teach the embedding a structure unrelated to the evaluation's ground
    truth and the reference's nDCG falls to chance)."""
    rng = np.random.default_rng(seed)
    zn = z / np.maximum(np.linalg.norm(z, axis=1, keepdims=True), 1e-12)
    zc = zn - zn.mean(0, keepdims=True)
    _, _, vt = np.linalg.svd(zc, full_matrices=False)
    xy = zc @ vt[:2].T
    xy = (xy - xy.min(0)) / np.maximum(xy.max(0) - xy.min(0), 1e-12)
    xy = (xy + rng.normal(0.0, jitter, size=xy.shape)) * extent
    d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    return np.exp(-d * 100.0)


def powerlaw_flow_graph(n_regions, L=24, seed=100000, mean_degree=42, cap=4096, zipf_a=1.6, pop_s=0.8):
    """Time-sliced synthetic flow graph: node (h, r) has id h*n_regions + r (first-appearance order of a
    host that walks layers then regions), edges (h, r) -> ((h+1)%L, r') grouped by source.

    out-degree ~ Zipf(zipf_a) rescaled to `mean_degree`, capped at `cap`; destinations by preferential
    attachment over regions: region popularity follows a rank-Zipf law, p(rank r) ~ (r+1)^-pop_s with pop_s = 0.8
    (the busiest 1 % of the regions attract ~40 % of the edges), assigned to region ids by a random permutation;
    weights floor(Pareto(1.2)) + 1.  (An unbounded Zipf *sample* as popularity lets one or two regions swallow the
    whole graph, which makes every walk and every embedding row cache-resident -- not the HBM-bound shape the
    100K / 1M-region configs are meant to exercise.)
    Returns dict(n_vertices, src, dst, w, sources, v_layer, v_region)."""
    rng = np.random.default_rng(seed)
    nv = n_regions * L
    raw = rng.zipf(zipf_a, size=nv).astype(np.float64)
    raw = np.minimum(raw, cap * 4.0)
    deg = np.maximum(1, np.minimum(cap, np.round(raw * (mean_degree / raw.mean())))).astype(np.int64)
    deg = np.minimum(deg, n_regions)
    ne = int(deg.sum())
    src = np.repeat(np.arange(nv, dtype=np.int32), deg)
    pop = (1.0 / np.arange(1, n_regions + 1, dtype=np.float64) ** pop_s)[rng.permutation(n_regions)]
    cdf = np.cumsum(pop / pop.sum())
    dst_region = np.searchsorted(cdf, rng.random(ne)).astype(np.int64)
    np.minimum(dst_region, n_regions - 1, out=dst_region)
    layer = (src // n_regions).astype(np.int64)
    dst = (((layer + 1) % L) * n_regions + dst_region).astype(np.int32)
    w = (np.floor(rng.pareto(1.2, size=ne)) + 1.0).astype(np.float64)
    v = np.arange(nv, dtype=np.int32)
    return dict(n_vertices=nv, src=src, dst=dst, w=w, sources=np.arange(n_regions, dtype=np.int32),
                v_layer=(v // n_regions).astype(np.int32), v_region=(v % n_regions).astype(np.int32))


def powerlaw_flow_graph_layered(n_regions, L=24, seed=1000000, mean_degree=42, cap=4096, zipf_a=1.6, pop_s=0.8, threads=None):
    """Same model as powerlaw_flow_graph, generated one layer at a time (one independent RNG stream and one host
    thread per layer) into preallocated arrays, so that the 1M-region x 24-slice config (~1.0e9 edges, 16 GB of COO)
    needs no multi-GB temporaries and a few tens of seconds on the host."""
    import concurrent.futures
    import os
    ss = np.random.SeedSequence(seed)
    rng = np.random.default_rng(ss.spawn(1)[0])
    nv = n_regions * L
    raw = np.minimum(rng.zipf(zipf_a, size=nv).astype(np.float64), cap * 4.0)
    deg = np.maximum(1, np.minimum(cap, np.round(raw * (mean_degree / raw.mean())))).astype(np.int64)
    deg = np.minimum(deg, n_regions)
    del raw
    first = np.concatenate([[0], np.cumsum(deg.reshape(L, n_regions).sum(axis=1))])
    ne = int(first[-1])
    src = np.empty(ne, np.int32)
    dst = np.empty(ne, np.int32)
    w = np.empty(ne, np.float64)
    pop = (1.0 / np.arange(1, n_regions + 1, dtype=np.float64) ** pop_s)[rng.permutation(n_regions)]
    cdf = np.cumsum(pop / pop.sum())
    streams = ss.spawn(L + 1)[1:]

    def layer(h):
        g = np.random.default_rng(streams[h])
        lo, hi = int(first[h]), int(first[h + 1])
        src[lo:hi] = np.repeat(np.arange(h * n_regions, (h + 1) * n_regions, dtype=np.int32), deg[h * n_regions:(h + 1) * n_regions])
        r = np.searchsorted(cdf, g.random(hi - lo))
        np.minimum(r, n_regions - 1, out=r)
        r += ((h + 1) % L) * n_regions
        dst[lo:hi] = r
        w[lo:hi] = np.floor(g.pareto(1.2, size=hi - lo)) + 1.0

    with concurrent.futures.ThreadPoolExecutor(max_workers=threads or min(L, os.cpu_count() or 1)) as ex:
        list(ex.map(layer, range(L)))
    return dict(n_vertices=nv, src=src, dst=dst, w=w, sources=np.arange(n_regions, dtype=np.int32))


def poi_latents():
    """Per-tract latent vectors: the 10 POI category counts of the reference's miscs/POI_tract.pickle
    (ground truth of python/embeddingEvaluation_tract.py:63-103), in tract_ids() order; zeros if absent."""
    header = ['Food', 'Residence', 'Travel', 'Arts & Entertainment', 'Outdoors & Recreation',
              'College & Education', 'Nightlife', 'Professional', 'Shops', 'Event']
    with open(os.path.join(_GOLDEN, "poi_tract.json")) as f:
        d = json.load(f)
    z = np.zeros((len(d["tract_ids"]), len(header)))
    for i, t in enumerate(d["tract_ids"]):
        for j, c in enumerate(header):
            z[i, j] = d["poi"].get(str(t), {}).get(c, 0)
    return z


def ca_latents():
    """Per-community-area latent vectors: the 20 binary label sets of miscs/{crime,lehd,demo,poi}-label."""
    with open(os.path.join(_GOLDEN, "ca_labels.json")) as f:
        d = json.load(f)
    cols = [d["crime-label"], d["lehd-label"]] + [d["demo-label"][k] for k in sorted(d["demo-label"])] + \
           [d["poi-label"][k] for k in sorted(d["poi-label"])]
    return np.array(cols, np.float64).T


def planted_flow_tensor(z, seed=2013, mean_trips_per_pair_hour=0.05, kappa=3.0):
    """F[src, hour, dst] int32 ~ Poisson(rate) with a planted structure, so that the reference's downstream
    metrics (which compare embeddings with POI / label ground truth) are informative on synthetic data:
        rate(s, h, d) ~ emit_s(h) * attract_d(h) * exp(kappa * cos(z_s, z_d))
    where emit / attract mix per-category hour profiles by the region's latent vector z (regions with similar
    z send and receive similar flows).  The real flow records are absent from the reference."""
    rng = np.random.default_rng(seed)
    n, c = z.shape
    zn = z / np.maximum(np.linalg.norm(z, axis=1, keepdims=True), 1e-12)
    hours = np.arange(24)
    prof_out = np.stack([np.exp(-0.5 * ((hours - rng.uniform(5, 22)) / rng.uniform(2, 5)) ** 2) for _ in range(c)])
    prof_in = np.stack([np.exp(-0.5 * ((hours - rng.uniform(5, 22)) / rng.uniform(2, 5)) ** 2) for _ in range(c)])
    size = 0.2 + z.sum(1) / max(z.sum(1).mean(), 1e-12)              # busier regions emit / attract more
    emit = (zn @ prof_out + 0.05) * size[:, None]                     # [n, 24]
    attract = (zn @ prof_in + 0.05) * size[:, None]                   # [n, 24]
    aff = np.exp(kappa * (zn @ zn.T))                                 # [n, n]
    F = np.empty((n, 24, n), np.int32)
    total_target = mean_trips_per_pair_hour * n * n * 24
    raw_total = sum(float((emit[:, h][:, None] * attract[:, h][None, :] * aff).sum()) for h in range(24))
    scale = total_target / raw_total
    for h in range(24):
        rate = emit[:, h][:, None] * attract[:, h][None, :] * aff * scale
        F[:, h, :] = rng.poisson(rate).astype(np.int32)
    return F
