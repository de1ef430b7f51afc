// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin C-ABI driver around the UNMODIFIED reference engine.  It is compiled
// together with /root/reference/src/fastsk/_fastsk/{fastsk_kernel,shared}.cpp
// (sources are read where they lie; nothing is copied into this repo) into
// oracle/_ref/libfskref.so by oracle/Makefile.  Recipe: SURVEY.md Appendix B.
//
// What it calls (all public in the reference):
//   extractFeatures                         shared.h:39 / shared.cpp:55-91
//   KernelFunction::kernel_build_parallel   fastsk_kernel.hpp:34 / fastsk_kernel.cpp:145-322
//   KernelFunction::compute_kernel          fastsk_kernel.hpp:33 / fastsk_kernel.cpp:24-106
//   KernelFunction::stdevs                  fastsk_kernel.hpp:31
//
// fskref_build drives kernel_build_parallel with a CALLER-SUPPLIED work queue
// (the reference shuffles with a wall-clock seed, fastsk_kernel.cpp:36-38), so
// the result is deterministic and the unnormalised K can be read before the
// normalisation loop (fastsk_kernel.cpp:96-103), which this file re-applies
// only on request.
#include "fastsk_kernel.hpp"
#include "shared.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <pthread.h>
#include <set>
#include <thread>
#include <vector>

namespace {

struct Prepared {
    std::vector<std::vector<int>> seqs;
    std::vector<int*> rows;
    std::vector<int> lengths;
    int dict_size = 0;
    int max_code = 0;
};

// Mirrors the argument marshalling of FastSK::compute_kernel (fastsk.cpp:30-93):
// row pointers, lengths, dict_size = |{0} U values|.
void prepare(const int32_t* codes, const int64_t* offsets, int64_t n, Prepared& p) {
    std::set<int> dict;
    dict.insert(0);
    p.seqs.resize(n);
    for (int64_t i = 0; i < n; ++i) {
        p.seqs[i].assign(codes + offsets[i], codes + offsets[i + 1]);
        for (int v : p.seqs[i]) {
            dict.insert(v);
            if (v > p.max_code) p.max_code = v;
        }
        p.rows.push_back(p.seqs[i].data());
        p.lengths.push_back((int)p.seqs[i].size());
    }
    p.dict_size = (int)dict.size();
}

void fill_params(kernel_params& kp, Features* f, const Prepared& p, int64_t n_train, int64_t n_test,
                 int g, int m, int T, int approx, double delta, int max_iters, int skip_variance) {
    long int total = n_train + n_test;
    kp.g = g;
    kp.k = g - m;
    kp.m = m;
    kp.n_str_train = n_train;
    kp.n_str_test = n_test;
    kp.total_str = total;
    kp.n_str_pairs = (total / (double)2) * (total + 1);   // fastsk.cpp:101
    kp.features = f;
    kp.dict_size = p.dict_size;
    kp.num_threads = T;
    kp.num_mutex = T;
    kp.workQueue = nullptr;
    kp.queueSize = 0;
    kp.quiet = true;
    kp.approx = approx != 0;
    kp.delta = delta;
    kp.max_iters = max_iters;
    kp.skip_variance = skip_variance != 0;
}

void normalise_packed(double* K, long int n) {   // same loop order as fastsk_kernel.cpp:96-103
    for (long int i = 0; i < n; i++)
        for (long int j = 0; j < i; j++)
            K[i * (i + 1) / 2 + j] = K[i * (i + 1) / 2 + j] /
                                     std::sqrt(K[i * (i + 1) / 2 + i] * K[j * (j + 1) / 2 + j]);
    for (long int i = 0; i < n; i++) {
        double d = K[i * (i + 1) / 2 + i];
        K[i * (i + 1) / 2 + i] = d / std::sqrt(d * d);
    }
}

}  // namespace

extern "C" {

// returns 0 on success; -1 codes not contiguous (reference would corrupt the heap);
// -2 N too large for the reference's int tri_access (shared.cpp:97-117); -3 g > shortest sequence.
int fskref_build(const int32_t* codes, const int64_t* offsets, int64_t n_train, int64_t n_test,
                 int g, int m, int T, int approx, double delta, int max_iters, int skip_variance,
                 const int32_t* queue, int nq, int normalise,
                 double* K_out, double* stdevs_out, int64_t stdevs_cap, int64_t* n_stdevs) {
    int64_t n = n_train + n_test;
    if (n > 46341) return -2;
    Prepared p;
    prepare(codes, offsets, n, p);
    if (p.max_code >= p.dict_size) return -1;
    for (int len : p.lengths) if (len < g) return -3;

    Features* f = extractFeatures(p.rows.data(), p.lengths, (int)n, g);
    if (T < 1) T = 20;                 // fastsk_kernel.cpp:54-60
    if (T > nq) T = nq;                // fastsk_kernel.cpp:61
    kernel_params kp;
    fill_params(kp, f, p, n_train, n_test, g, m, T, approx, delta, max_iters, skip_variance);

    std::vector<WorkItem> wq(nq);
    for (int i = 0; i < nq; ++i) { wq[i].m = m; wq[i].combo_num = queue[i]; }
    std::vector<pthread_mutex_t> mtx(T);
    for (auto& mu : mtx) pthread_mutex_init(&mu, NULL);
    std::memset(K_out, 0, sizeof(double) * (size_t)kp.n_str_pairs);

    KernelFunction kf(&kp);
    std::vector<std::thread> th;
    for (int tid = 0; tid < T; ++tid)
        th.emplace_back(&KernelFunction::kernel_build_parallel, &kf, tid, wq.data(), nq, mtx.data(), &kp, K_out);
    for (auto& t : th) t.join();
    for (auto& mu : mtx) pthread_mutex_destroy(&mu);

    if (normalise) normalise_packed(K_out, (long int)n);
    if (n_stdevs) {
        *n_stdevs = (int64_t)kf.stdevs.size();
        for (int64_t i = 0; i < *n_stdevs && i < stdevs_cap; ++i) stdevs_out[i] = kf.stdevs[i];
    }
    free(f->features);
    free(f->group);
    free(f);
    return 0;
}

// The reference's own top-level engine call (shuffle seeded by wall clock, thread fan-out,
// merge, normalisation): used only as the timed CPU baseline (bench.py --impl reference).
int fskref_compute_kernel(const int32_t* codes, const int64_t* offsets, int64_t n_train, int64_t n_test,
                          int g, int m, int T, int approx, double delta, int max_iters, int skip_variance,
                          double* K_out /* may be NULL */) {
    int64_t n = n_train + n_test;
    if (n > 46341) return -2;
    Prepared p;
    prepare(codes, offsets, n, p);
    if (p.max_code >= p.dict_size) return -1;
    for (int len : p.lengths) if (len < g) return -3;
    Features* f = extractFeatures(p.rows.data(), p.lengths, (int)n, g);
    kernel_params kp;
    fill_params(kp, f, p, n_train, n_test, g, m, T, approx, delta, max_iters, skip_variance);
    kp.num_mutex = -1;
    KernelFunction kf(&kp);
    double* K = kf.compute_kernel();
    if (K_out) std::memcpy(K_out, K, sizeof(double) * (size_t)kp.n_str_pairs);
    free(K);
    free(f->features);
    free(f->group);
    free(f);
    return 0;
}

}  // extern "C"
