#!/usr/bin/env python
"""The reference's quick validation (test/run_check.py) on this backend: gapped k-mer kernel of the bundled EP300 TFBS
set with FastSK(g=10, m=6, t=1, approx=True), a linear SVM on the kernel rows, and the reference's only assertion,
AUC >= 0.9.  Only the import differs: `from fastsk_b200 import FastSK, FastaUtility`.

    python examples/run_check.py [--train data/EP300.train.fasta --test data/EP300.test.fasta]
"""
import argparse
import os
import sys
import time

import numpy as np
from sklearn.calibration import CalibratedClassifierCV
from sklearn.metrics import roc_auc_score
from sklearn.svm import LinearSVC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastsk_b200 import FastSK, FastaUtility  # noqa: E402


def run(train, test, **fastsk_args):
    reader = FastaUtility()
    Xtrain, Ytrain = reader.read_data(train)
    Xtest, Ytest = reader.read_data(test)
    Ytest = np.array(Ytest).reshape(-1, 1)

    t0 = time.perf_counter()
    fastsk = FastSK(**(fastsk_args or dict(g=10, m=6, t=1, approx=True)))
    fastsk.compute_kernel(Xtrain, Xtest)
    Ktrain = fastsk.get_train_kernel()
    Ktest = fastsk.get_test_kernel()
    kernel_s = time.perf_counter() - t0

    svm = LinearSVC(C=1)
    clf = CalibratedClassifierCV(svm, cv=5).fit(Ktrain, Ytrain)
    acc = clf.score(Ktest, Ytest)
    auc = roc_auc_score(Ytest, clf.predict_proba(Ktest)[:, 1])
    return acc, auc, kernel_s, len(fastsk.get_stdevs())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--train", default=os.path.join(ROOT, "data", "EP300.train.fasta"))
    ap.add_argument("--test", default=os.path.join(ROOT, "data", "EP300.test.fasta"))
    a = ap.parse_args()
    acc, auc, kernel_s, iters = run(a.train, a.test)
    print("Linear SVM:\n\tAcc = {}, AUC = {}  (kernel in {:.3f} s, {} sampled combinations)".format(acc, auc, kernel_s, iters))
    assert auc >= 0.9, "AUC is not correct. Should be >= 0.9. Received: {}".format(auc)
