// TEST-ONLY host evaluator of the feature program (never linked into libkoopfit.so).
// It compiles the same program.cpp / lift_eval.h the CUDA library uses, so that the
// dictionary compiler and op semantics can be checked against the oracle on a box
// without a GPU.  Build: g++ -O2 -ffp-contract=off -shared -fPIC (see tests/conftest.py).
#include <cstring>
#include <string>
#include <vector>

#include "../koopman-realizations_b200/csrc/lift_eval.h"
#include "../koopman-realizations_b200/csrc/program.h"

extern "C" int hostlift_dims(const kf_basis* basis, int* n_full, int* N) {
    KfProgram prog;
    std::string err;
    int rc = kf_build_program(basis, prog, err);
    if (rc) return rc;
    *n_full = prog.n_full();
    *N = prog.N();
    return 0;
}

// ops out: 4 ints (kind,a,b,0) + 1 double per feature
extern "C" int hostlift_ops(const kf_basis* basis, int* kinds, int* as, int* bs, double* cs) {
    KfProgram prog;
    std::string err;
    int rc = kf_build_program(basis, prog, err);
    if (rc) return rc;
    for (int j = 0; j < prog.n_full(); ++j) {
        kinds[j] = prog.ops[j].kind; as[j] = prog.ops[j].a; bs[j] = prog.ops[j].b; cs[j] = prog.ops[j].c;
    }
    return 0;
}

// V: rows x nv column-major; out: rows x n_full column-major (full dictionary, no pcs)
extern "C" int hostlift_full(const kf_basis* basis, long long rows, const double* V, double* out) {
    KfProgram prog;
    std::string err;
    int rc = kf_build_program(basis, prog, err);
    if (rc) return rc;
    const int nf = prog.n_full(), nv = prog.nv;
    std::vector<double> f(nf);
    for (long long s = 0; s < rows; ++s) {
        for (int i = 0; i < nv; ++i) f[i] = V[(size_t)i * rows + s];
        for (int j = nv; j < nf; ++j)
            f[j] = kf_eval_op(prog.ops[j], nv, prog.centres.data(), [&](int k) { return f[k]; });
        for (int j = 0; j < nf; ++j) out[(size_t)j * rows + s] = f[j];
    }
    return 0;
}
