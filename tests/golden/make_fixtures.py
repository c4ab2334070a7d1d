"""Generates the committed fixtures under tests/golden/ from the reference checkout.

Run HERE (the container that has /root/reference); the GPU box only sees the committed outputs.
    python tests/golden/make_fixtures.py

Outputs
  layered_graph_test.json  the reference's only hot-path golden vector, transcribed from
                           embedding/src/test/java/embedding/LayeredGraphTest.java:12-44
  poi_tract.json           the 801 sorted tract ids and the POI category counts of
                           miscs/POI_tract.pickle (ground truth of python/embeddingEvaluation_tract.py:63-103)
  ca_labels.json           miscs/crime-label, lehd-label, demo-label, poi-label (binary labels of the 77
                           community areas used by python/binaryClassification_CA.py:40-46)
  published.json           the reference's published walk timings (python/running_time.py:16-20) and nDCG
                           table (python/ndcg.pickle)
"""
import json
import os
import pickle
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    # --- golden vector: parsed from the Java test source so the numbers are the reference's own
    src = open(os.path.join(REF, "embedding/src/test/java/embedding/LayeredGraphTest.java")).read()
    weights = [float(x) for x in re.findall(r"new LayeredGraph\.Edge\(org, d\d, (\d+)\)", src)]
    alias = [int(x) for x in re.findall(r"assertEquals\(org\.aliasTable\[\d\], (-?\d+)\)", src)]
    prob = [float(x) for x in re.findall(r"assertEquals\(org\.probTable\[\d\], ([\d.]+)\)", src)]
    out_degree = float(re.search(r"assertEquals\(org\.outDegree, ([\d.]+)\)", src).group(1))
    samples = [(float(x), int(i)) for x, i in re.findall(r"sampleNextVertex\(([\d.]+)\)\.id, (\d)\)", src)]
    assert weights == [2.0, 10.0, 8.0] and len(alias) == 3 and len(prob) == 3 and len(samples) == 5
    json.dump(dict(source="LayeredGraphTest.java:12-44", weights=weights, alias=alias, prob=prob,
                   out_degree=out_degree, samples=samples),
              open(os.path.join(OUT, "layered_graph_test.json"), "w"), indent=1)

    with open(os.path.join(REF, "miscs/POI_tract.pickle"), "rb") as f:
        ids = pickle.load(f, encoding="latin1")
        poi = pickle.load(f, encoding="latin1")
    json.dump(dict(source="miscs/POI_tract.pickle", tract_ids=[int(i) for i in ids],
                   poi={str(int(k)): {c: int(n) for c, n in v.items()} for k, v in poi.items()}),
              open(os.path.join(OUT, "poi_tract.json"), "w"))

    labels = {}
    for name in ("crime-label", "lehd-label", "demo-label", "poi-label"):
        o = pickle.load(open(os.path.join(REF, "miscs", name), "rb"), encoding="latin1")
        labels[name] = {k: [int(x) for x in v] for k, v in o.items()} if isinstance(o, dict) else [int(x) for x in o]
    json.dump(dict(source="miscs/{crime,lehd,demo,poi}-label", **labels),
              open(os.path.join(OUT, "ca_labels.json"), "w"))

    rt = open(os.path.join(REF, "python/running_time.py")).read()
    n_seqs = [float(x) for x in re.search(r"n_seqs = \[([\d., ]+)\]", rt).group(1).split(",")]
    block = rt[rt.index("running_time = "):rt.index("lines_setting")]
    times = [[float(x) for x in row.split(",")] for row in re.findall(r"\[([\d., ]+)\]", block)]
    names = re.findall(r'"([^"]+)"', rt[rt.index("lines_setting"):rt.index("lines_style")])
    assert len(times) == 4 and len(names) == 4
    ndcg = pickle.load(open(os.path.join(REF, "python/ndcg.pickle"), "rb"), encoding="latin1")
    json.dump(dict(source="python/running_time.py:16-20, python/ndcg.pickle",
                   million_walks=n_seqs, seconds=dict(zip(names, times)),
                   ndcg=[[float(x) for x in r] for r in ndcg]),
              open(os.path.join(OUT, "published.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
