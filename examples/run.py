"""Command-line demo of the kernel function as an empirical kernel map (EKM) for a linear SVM -- the same flags as the
reference's examples/run.py:23-72 (--trn --tst -g -m -C -t -a -d -I --skip-variance), on the B200 library.

    python examples/run.py --trn data/EP300.train.fasta --tst data/EP300.test.fasta -g 10 -m 6 [-a -I 50] [--gpu-svm]

--gpu-svm trains the linear SVM on the device-resident kernel (FastSK.fit_linear_gpu: no 16 GB device -> host copy at
N = 50 000); without it the script does what the reference does: scikit-learn's LinearSVC + CalibratedClassifierCV on host
copies of the two kernels."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastsk_b200 import FastSK, FastaUtility  # noqa: E402


def get_args():
    p = argparse.ArgumentParser(description="fastsk_b200 evaluation (flags of the reference's examples/run.py)")
    p.add_argument("--trn", type=str, required=True, help="Training file", metavar="1.1.train.fasta")
    p.add_argument("--tst", type=str, required=True, help="Test file", metavar="1.1.test.fasta")
    p.add_argument("-g", type=int, required=True)
    p.add_argument("-m", type=int, required=True)
    p.add_argument("-C", type=float, default=1)
    p.add_argument("-t", type=int, default=20, help="Number of virtual streams of the approximation algorithm (the reference's threads)")
    p.add_argument("-a", "--approx", action="store_true", default=False, help="Flag to enable the approximation algorithm")
    p.add_argument("-d", "--delta", type=float, default=0.025, help="Delta parameter for approximation algorithm")
    p.add_argument("-I", type=int, default=50, help="Maximum number of iterations to use if running the approximation algorithm")
    p.add_argument("--skip-variance", action="store_true", default=False,
                   help="Can use approximation algo for the given iterations, but will not compute variances")
    p.add_argument("--seed", type=int, default=None, help="seed of the combination shuffle (the reference uses the wall clock)")
    p.add_argument("--gpu-svm", action="store_true", help="train the linear SVM on the device-resident kernel")
    return p.parse_args()


def main():
    args = get_args()
    reader = FastaUtility()
    Xtrain, Ytrain = reader.read_data(args.trn)
    Xtest, Ytest = reader.read_data(args.tst)
    start = time.time()
    fastsk = FastSK(g=args.g, m=args.m, t=args.t, approx=args.approx, max_iters=args.I, delta=args.delta,
                    skip_variance=args.skip_variance, seed=args.seed)
    fastsk.compute_kernel(Xtrain, Xtest)
    print("Kernel computation time: ", time.time() - start)
    if args.gpu_svm:
        t0 = time.time()
        fastsk.fit_linear_gpu(Ytrain, C=args.C)
        acc, auc = fastsk.score_gpu(Ytest, "accuracy") / 100.0, fastsk.score_gpu(Ytest, "auc")
        print("Linear SVM on the device ({:.3f} s):\n\tAcc = {}, AUC = {}".format(time.time() - t0, acc, auc))
        return
    from sklearn import metrics
    from sklearn.calibration import CalibratedClassifierCV
    from sklearn.svm import LinearSVC
    Ktr, Kte = fastsk.get_train_kernel(), fastsk.get_test_kernel()
    clf = CalibratedClassifierCV(LinearSVC(C=args.C), cv=5).fit(Ktr, Ytrain)
    acc = clf.score(Kte, np.array(Ytest).reshape(-1, 1))
    auc = metrics.roc_auc_score(Ytest, clf.predict_proba(Kte)[:, 1])
    print("Linear SVM:\n\tAcc = {}, AUC = {}".format(acc, auc))


if __name__ == "__main__":
    main()
