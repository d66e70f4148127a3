// Evaluation of one feature-program op.  Shared by the CUDA lift kernels (device) and by
// the test-only host evaluator (tests/hostlift.cpp, compiled with -ffp-contract=off) so
// that the op semantics are checked against the oracle without a GPU.
//
// F(j) reads feature j of the current snapshot; every feature with index < j has been
// stored already (MUL only references earlier features).  Products/sums that define the
// rounding are written with KF_MUL / KF_ADD / KF_SUB which never fuse.
#pragma once
#include "program.h"

#if defined(__CUDA_ARCH__)
#define KF_HD __host__ __device__ __forceinline__
#define KF_MUL(x, y) __dmul_rn((x), (y))
#define KF_ADD(x, y) __dadd_rn((x), (y))
#define KF_SUB(x, y) __dsub_rn((x), (y))
#elif defined(__CUDACC__)
#define KF_HD __host__ __device__ __forceinline__
#define KF_MUL(x, y) ((x) * (y))
#define KF_ADD(x, y) ((x) + (y))
#define KF_SUB(x, y) ((x) - (y))
#else
#include <cmath>
#define KF_HD inline
#define KF_MUL(x, y) ((x) * (y))
#define KF_ADD(x, y) ((x) + (y))
#define KF_SUB(x, y) ((x) - (y))
#endif

// physicists' Hermite polynomial by H_{j+1} = (2x) H_j - (2j) H_{j-1}  (hermiteH, Ksysid.m:827-829)
KF_HD double kf_hermite(int k, double x) {
    if (k == 0) return 1.0;
    const double tx = KF_MUL(2.0, x);
    double h0 = 1.0, h1 = tx;
    for (int j = 1; j < k; ++j) {
        const double h2 = KF_SUB(KF_MUL(tx, h1), KF_MUL(2.0 * (double)j, h0));
        h0 = h1;
        h1 = h2;
    }
    return h1;
}

// Feat: callable double(int j) returning an earlier feature (features 0..nv-1 are v itself).
template <class Feat>
KF_HD double kf_eval_op(const KfOp& op, int nv, const double* centres, Feat F) {
    switch (op.kind) {
        case KF_OP_VAR:
            return F(op.a);   // the caller materialises v into features 0..nv-1 first
        case KF_OP_CONST:
            return op.c;
        case KF_OP_MUL:
            return KF_MUL(F(op.a), F(op.b));
        case KF_OP_COS:
            return cos(KF_MUL(op.c, F(op.a)));
        case KF_OP_SIN:
            return sin(KF_MUL(op.c, F(op.a)));
        case KF_OP_HERM:
            return kf_hermite(op.b, F(op.a));
        case KF_OP_GAUSS: {   // exp(-||v - c||^2), Ksysid.m:805-806
            double acc = 0.0;
            const double* c = centres + (size_t)op.a * nv;
            for (int i = 0; i < nv; ++i) {
                const double d = KF_SUB(F(i), c[i]);
                acc = KF_ADD(acc, KF_MUL(d, d));
            }
            return exp(-acc);
        }
        default:
            return 0.0;
    }
}
