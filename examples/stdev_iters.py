#!/usr/bin/env python
"""Standard deviation of the approximated kernel and test AUC against the number of iterations -- the reference's
`results/run_experiments.py --stdev-I` experiment (stdev_and_auc_vs_iters_experiments, :1098-1191) on this backend.

For I in the reference's schedule (1, 2, 4, 6, 8, 10, 15, ..., 45, 50, 70, ...) and five differently-seeded shuffles it runs
FastSK(g, m, t=1, approx=True, max_iters=I, delta=0.025), records get_stdevs()[-1] and trains / evaluates the linear SVM on
the empirical kernel map.  Writes the reference's CSV columns; with --compare it prints the mean stdev per I next to the
numbers the reference published for EP300 g=10 m=4 (results/spreadsheets/stdevs/EP300_stdev_auc_iters.csv:2-12).

    python examples/stdev_iters.py --dataset EP300 -g 10 -m 4 [--max-I 50] [--gpu-svm] [--compare]
"""
import argparse
import csv
import os
import sys
from math import comb

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastsk_b200 import FastSK, FastaUtility  # noqa: E402

# mean stdev over five samples, reference repo, EP300 g=10 k=6 m=4 C=1 (EP300_stdev_auc_iters.csv:2-12, column "mean stdev")
REFERENCE_EP300_G10_M4 = {1: 3162.2775020544923, 2: 1.0780596600419674, 4: 0.4945835896129476, 6: 0.2929101502865293,
                          8: 0.21269915979077686, 10: 0.1660921390593993, 15: 0.11375112626424948, 20: 0.0864592240196981,
                          25: 0.0656943400983907, 30: 0.05600643808403497, 35: 0.04972645977235169}


def schedule(max_I):
    """run_experiments.py:1134-1147"""
    iters = [1]
    if max_I > 1:
        iters += list(range(2, min(max_I, 10), 2))
        if max_I >= 10:
            iters += list(range(10, min(max_I, 50), 5))
        if max_I >= 50:
            iters += list(range(50, max_I, 20))
        iters += [max_I]
    return iters


def ci(x, z=2.776):     # 95 % two-sided t interval for five samples (get_CI of the reference's utils)
    x = np.asarray(x, dtype=np.float64)
    h = z * x.std(ddof=1) / np.sqrt(len(x)) if len(x) > 1 else 0.0
    return float(x.mean()), float(x.mean() - h), float(x.mean() + h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="EP300")
    ap.add_argument("-g", type=int, default=10)
    ap.add_argument("-m", type=int, default=4)
    ap.add_argument("-C", type=float, default=1.0)
    ap.add_argument("--max-I", type=int, default=500)
    ap.add_argument("--samples", type=int, default=5)
    ap.add_argument("--gpu-svm", action="store_true", help="train the SVM on the device-resident kernel (fit_linear_gpu)")
    ap.add_argument("--compare", action="store_true")
    ap.add_argument("--output-dir", default=".")
    a = ap.parse_args()
    fu = FastaUtility()
    Xtr, Ytr = fu.read_data(os.path.join(ROOT, "data", a.dataset + ".train.fasta"))
    Xte, Yte = fu.read_data(os.path.join(ROOT, "data", a.dataset + ".test.fasta"))
    max_I = min(comb(a.g, a.m), 500, a.max_I)
    rows = []
    for I in schedule(max_I):
        accs, aucs, sds = [], [], []
        for s in range(a.samples):
            f = FastSK(g=a.g, m=a.m, t=1, approx=True, max_iters=I, delta=0.025, seed=1000 * I + s)
            f.compute_kernel(Xtr, Xte)
            sd = f.get_stdevs()
            assert len(sd) == I or a.g - a.m == 0, "the stream converged before max_iters"
            if a.gpu_svm:
                f.fit_linear_gpu(Ytr, C=a.C)
                acc, auc = f.score_gpu(Yte, "accuracy") / 100.0, f.score_gpu(Yte, "auc")
            else:
                from sklearn import metrics
                from sklearn.calibration import CalibratedClassifierCV
                from sklearn.svm import LinearSVC
                clf = CalibratedClassifierCV(LinearSVC(C=a.C, class_weight="balanced"), cv=5).fit(f.get_train_kernel(), Ytr)
                Kte = f.get_test_kernel()
                acc, auc = clf.score(Kte, Yte), metrics.roc_auc_score(Yte, clf.predict_proba(Kte)[:, 1])
            accs.append(acc); aucs.append(auc); sds.append(sd[-1])
        row = {"dataset": a.dataset, "g": a.g, "k": a.g - a.m, "m": a.m, "C": a.C, "iters": I}
        for name, vals in (("acc", accs), ("auc", aucs), ("stdev", sds)):
            for i, v in enumerate(vals):
                row[f"{name} sample {i + 1}"] = v
            row[f"mean {name}"], row[f"lower {name}"], row[f"upper {name}"] = ci(vals)
        rows.append(row)
        ref = REFERENCE_EP300_G10_M4.get(I) if (a.compare and (a.dataset, a.g, a.m) == ("EP300", 10, 4)) else None
        print(f"I = {I:4d}  mean stdev = {row['mean stdev']:.6g}  mean auc = {row['mean auc']:.6f}" +
              (f"  reference mean stdev = {ref:.6g}  ratio = {row['mean stdev'] / ref:.3f}" if ref else ""), flush=True)
    out = os.path.join(a.output_dir, a.dataset + "_stdev_auc_iters.csv")
    with open(out, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=list(rows[0]))
        w.writeheader()
        w.writerows(rows)
    print("wrote", out)


if __name__ == "__main__":
    main()
