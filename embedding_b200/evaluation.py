"""Downstream metrics of the reference, ported to Python 3 / numpy (CPU; they are the parity yardstick for
stage 2, not part of the accelerated path -- SURVEY.md 8(c)):

  * tract level: pairwise-similarity nDCG@k against the POI ground truth
        python/embeddingEvaluation_tract.py:63-103 (generatePairWiseGT), :139-166 (.vec parser),
        :169-196 (pairwiseEstimator), :249-260 (dcg_atK / ndcg_atK), :285-367 (evalute_by_pairwise_similarity)
  * CA level: 10-fold cross-validated DecisionTree / SVC accuracy on binary labels
        python/binaryClassification_CA.py:33-58
"""
import numpy as np


def read_vec(path):
    """retrieveCrossIntervalEmbeddings(fn, skipheader=0) :139-166 -> {layer: (features [m, D], region ids [m])}."""
    feats, rids = {}, {}
    with open(path) as f:
        for line in f:
            parts = line.rstrip("\n").split(" ")
            if len(parts) < 2:
                continue
            k1, k2 = parts[0].split("-")
            feats.setdefault(int(k1), []).append([float(x) for x in parts[1:]])
            rids.setdefault(int(k1), []).append(int(k2))
    return {k: (np.array(feats[k]), np.array(rids[k])) for k in feats}


def layers_from_model(syn0, id_of_word, v_layer, v_region):
    """Same structure as read_vec() straight from a trained model (rows in vocabulary order = .vec line order)."""
    out = {}
    lay, reg = np.asarray(v_layer)[id_of_word], np.asarray(v_region)[id_of_word]
    for h in np.unique(lay):
        m = lay == h
        out[int(h)] = (np.asarray(syn0)[m], reg[m])
    return out


def cosine_distance_matrix(X):
    """scipy.spatial.distance.cosine for all pairs; NaN (zero vector) -> 2 as the reference does (:99-100, :186-189)."""
    X = np.asarray(X, np.float64)
    norm = np.linalg.norm(X, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        D = 1.0 - (X @ X.T) / (norm[:, None] * norm[None, :])
    D[~np.isfinite(D)] = 2.0
    return D


class PairwiseGroundTruth:
    """generatePairWiseGT :63-103 over the ordered tract ids and their POI count vectors."""

    def __init__(self, ord_key, poi_vectors):
        self.ids = np.asarray(ord_key)
        self.index = {int(r): i for i, r in enumerate(self.ids)}
        self.D = cosine_distance_matrix(poi_vectors)
        n = len(self.ids)
        Dm = self.D.copy()
        Dm[np.arange(n), np.arange(n)] = np.inf          # `if k2 == k: continue`
        self.order = np.argsort(Dm, axis=1, kind="stable")[:, :n - 1]

    def dcg_max(self, topk):
        """dcg_atK(topk, gnd_est[rid], pair_gnd[rid]) :302-304."""
        top = self.order[:, :topk]
        relv = 1.0 - np.take_along_axis(self.D, top, axis=1)
        return (relv / np.log2(np.arange(2, topk + 2))[None, :]).sum(1)


def ndcg_at_k(gt, features, rids, topk):
    """pairwiseEstimator :169-196 + ndcg_atK :254-260 for one layer; test set = every region of the layer."""
    rids = np.asarray(rids)
    m = len(rids)
    if m <= topk:
        return float("nan")
    gi = np.array([gt.index[int(r)] for r in rids])
    D = cosine_distance_matrix(features)
    D[np.arange(m), np.arange(m)] = np.inf
    nb = np.argsort(D, axis=1, kind="stable")[:, :topk]           # neighbours, local indices
    relv = 1.0 - gt.D[gi[:, None], gi[nb]]
    dcg = (relv / np.log2(np.arange(2, topk + 2))[None, :]).sum(1)
    return float(np.mean(dcg / gt.dcg_max(topk)[gi]))


def pairwise_ndcg(gt, layers, ks=(5, 10, 20, 30, 40, 50, 60, 70)):
    """evalute_by_pairwise_similarity :285-367 for one embedding: mean over layers of nDCG@k, for each k (:790)."""
    out = {}
    for k in ks:
        vals = [ndcg_at_k(gt, f, r, k) for _, (f, r) in sorted(layers.items())]
        vals = [v for v in vals if np.isfinite(v)]
        out[int(k)] = float(np.mean(vals)) if vals else float("nan")
    return out


def pairwise_ndcg_device(ctx, gt, layers, ks=(5, 10, 20, 30, 40, 50, 60, 70)):
    """pairwise_ndcg() with the kNN + DCG of every layer computed by libdge (dge_eval_ndcg)."""
    out = {}
    for k in ks:
        vals = []
        for _, (f, r) in sorted(layers.items()):
            if len(r) <= k:
                continue
            gi = np.array([gt.index[int(x)] for x in r], np.int32)
            vals.append(ctx.eval_ndcg(f, gi, gt.D, k)[1])
        out[int(k)] = float(np.mean(vals)) if vals else float("nan")
    return out


def knn_overlap(layers_a, layers_b, k=10):
    """Mean Jaccard-free overlap |kNN_a(r) & kNN_b(r)| / k over regions and layers: how much two embeddings agree
    on neighbourhoods (an extra, label-free parity measure; not in the reference)."""
    vals = []
    for h in sorted(set(layers_a) & set(layers_b)):
        fa, ra = layers_a[h]
        fb, rb = layers_b[h]
        common, ia, ib = np.intersect1d(ra, rb, return_indices=True)
        if len(common) <= k:
            continue
        Da, Db = cosine_distance_matrix(fa[ia]), cosine_distance_matrix(fb[ib])
        m = len(common)
        Da[np.arange(m), np.arange(m)] = np.inf
        Db[np.arange(m), np.arange(m)] = np.inf
        na = np.argsort(Da, axis=1, kind="stable")[:, :k]
        nbb = np.argsort(Db, axis=1, kind="stable")[:, :k]
        hit = [(len(np.intersect1d(x, y)) / k) for x, y in zip(na, nbb)]
        vals.append(np.mean(hit))
    return float(np.mean(vals)) if vals else float("nan")


def knn_table(layers, region_ids, n_layers, k=10):
    """The k nearest neighbours (cosine, as pairwiseEstimator :169-196) of every region in every layer as one array
    [n_layers, n_regions, k] of positions in `region_ids` (int16; -1 = the region has no row in that layer): a compact
    fingerprint of an embedding's neighbourhood structure (committed for the oracle, tests/golden/)."""
    pos = {int(r): i for i, r in enumerate(region_ids)}
    out = np.full((n_layers, len(region_ids), k), -1, np.int16)
    for h, (f, r) in layers.items():
        m = len(r)
        if m <= k or not (0 <= h < n_layers):
            continue
        idx = np.array([pos[int(x)] for x in r])
        D = cosine_distance_matrix(f)
        D[np.arange(m), np.arange(m)] = np.inf
        nb = np.argsort(D, axis=1, kind="stable")[:, :k]
        out[h, idx] = idx[nb]
    return out


def knn_table_overlap(a, b):
    """Mean |kNN_a(r) & kNN_b(r)| / k over the (layer, region) rows present in both tables."""
    both = (a[:, :, 0] >= 0) & (b[:, :, 0] >= 0)
    A, B = a[both], b[both]
    k = A.shape[1]
    hits = (A[:, :, None] == B[:, None, :]).any(-1).sum(-1)
    return float(hits.mean() / k) if len(A) else float("nan")


def ca_classification_accuracy(layers, labels, region_ids, cv=10):
    """binaryClassification_CA.py:33-58 with the per-layer embeddings as features: mean 10-fold accuracy of a
    decision tree and an SVC over layers and label sets (labels: {name: [77 binary]})."""
    from sklearn import svm, tree
    from sklearn.model_selection import cross_val_score
    pos = {int(r): i for i, r in enumerate(region_ids)}
    res = {"DT": [], "SVM": []}
    for h, (f, r) in sorted(layers.items()):
        idx = np.array([pos[int(x)] for x in r])
        for name, lab in labels.items():
            y = np.asarray(lab)[idx]
            if y.sum() in (0, len(y)) or min(y.sum(), len(y) - y.sum()) < cv:
                continue
            res["DT"].append(cross_val_score(tree.DecisionTreeClassifier(random_state=0), f, y, cv=cv).mean())
            res["SVM"].append(cross_val_score(svm.SVC(), f, y, cv=cv).mean())
    return {k: float(np.mean(v)) if v else float("nan") for k, v in res.items()}
