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

// The lift evaluated GROUP by group exactly as kf_lift_tile_kernel does (slots, level order, store lists), with the
// invariants checked on the way: an op only reads slots written by the variables or an EARLIER level of its own group,
// every group stays within max_slots (unless it is a single feature's closure), every feature is stored exactly once.
// Returns 0, or a negative code naming the violated invariant.  out: rows x n_full column-major.
extern "C" int hostlift_groups(const kf_basis* basis, int max_slots, int single, long long rows, const double* V, double* out,
                               int* ngroups, int* max_used) {
    KfProgram prog;
    std::string err;
    int rc = kf_build_program(basis, prog, err);
    if (rc) return rc;
    std::vector<LtOp> gops;
    std::vector<LtStore> gstore;
    std::vector<LtGroup> groups;
    if (!kf_build_lift_groups(prog, max_slots, single != 0, gops, gstore, groups)) return -1;
    const int nf = prog.n_full(), nv = prog.nv;
    std::vector<int> stored(nf, 0);
    *ngroups = (int)groups.size();
    *max_used = 0;
    for (const LtGroup& g : groups) {
        if (g.nslots > *max_used) *max_used = g.nslots;
        std::vector<int> wlevel(g.nslots, 1 << 30);          // level at which a slot becomes readable
        for (int v = 0; v < nv; ++v) wlevel[v] = -1;
        for (int l = 0; l < g.nlevels; ++l)
            for (int e = g.level_start[l]; e < g.level_start[l + 1]; ++e) {
                const LtOp& op = gops[g.op_off + e];
                if (op.j < nv || op.j >= g.nslots) return -2;
                if (op.kind == KF_OP_MUL && (wlevel[op.a] >= l || wlevel[op.b] >= l)) return -3;
                if ((op.kind == KF_OP_COS || op.kind == KF_OP_SIN || op.kind == KF_OP_HERM || op.kind == KF_OP_VAR) && op.a >= nv) return -4;
                wlevel[op.j] = l;
            }
        if (g.level_start[g.nlevels] != g.nops) return -5;
        for (int e = 0; e < g.nst; ++e) {
            const LtStore& sr = gstore[g.st_off + e];
            if (sr.slot < 0 || sr.slot >= g.nslots || sr.row < 0 || sr.row >= nf || wlevel[sr.slot] == (1 << 30)) return -6;
            if (sr.slot != g.slot0 + e || sr.row != g.row0 + e) return -8;      // one contiguous slot range <-> one contiguous row range
            stored[sr.row] += 1;
        }
        std::vector<double> sh(g.nslots);
        for (long long s = 0; s < rows; ++s) {
            for (int i = 0; i < nv; ++i) sh[i] = V[(size_t)i * rows + s];
            for (int e = 0; e < g.nops; ++e) {
                const LtOp& op = gops[g.op_off + e];
                KfOp o{};
                o.kind = op.kind; o.a = op.a; o.b = op.b; o.c = op.c;
                sh[op.j] = kf_eval_op(o, nv, prog.centres.data(), [&](int k) { return sh[k]; });
            }
            for (int e = 0; e < g.nst; ++e) out[(size_t)gstore[g.st_off + e].row * rows + s] = sh[gstore[g.st_off + e].slot];
        }
    }
    for (int j = 0; j < nf; ++j)
        if (stored[j] != 1) return -7;
    return 0;
}

// The lift evaluated by ROW GROUPS exactly as kf_lift_stream_kernel does: slot ops level by level, then one row op per output
// row.  Invariants: operands are variables or slots written by an earlier level; the groups tile [0, n_full) in order; the
// slot / row budgets hold.  out: rows x n_full column-major.
extern "C" int hostlift_rowgroups(const kf_basis* basis, int max_slots, int max_rows, long long rows, const double* V, double* out,
                                  int* ngroups, int* max_used) {
    KfProgram prog;
    std::string err;
    int rc = kf_build_program(basis, prog, err);
    if (rc) return rc;
    std::vector<LtOp> gops;
    std::vector<LtGroup> groups;
    if (!kf_build_lift_rowgroups(prog, max_slots, max_rows, gops, groups)) return -1;
    const int nf = prog.n_full(), nv = prog.nv;
    *ngroups = (int)groups.size();
    *max_used = 0;
    int next_row = 0;
    for (const LtGroup& g : groups) {
        if (g.nslots > *max_used) *max_used = g.nslots;
        if (g.row0 != next_row || g.nst <= 0) return -2;
        if (max_rows > 0 && g.nst > max_rows) return -9;
        next_row += g.nst;
        std::vector<int> wlevel(g.nslots, 1 << 30);
        for (int v = 0; v < nv; ++v) wlevel[v] = -1;
        for (int l = 0; l < g.nlevels; ++l)
            for (int e = g.level_start[l]; e < g.level_start[l + 1]; ++e) {
                const LtOp& op = gops[g.op_off + e];
                if (op.j < nv || op.j >= g.nslots || wlevel[op.j] != (1 << 30)) return -3;
                if (op.kind == KF_OP_MUL && (wlevel[op.a] >= l || wlevel[op.b] >= l)) return -4;
                if ((op.kind == KF_OP_COS || op.kind == KF_OP_SIN || op.kind == KF_OP_HERM) && op.a >= nv) return -5;
                wlevel[op.j] = l;
            }
        if (g.level_start[g.nlevels] != g.nops || g.rop_off != g.op_off + g.nops) return -6;
        for (int e = 0; e < g.nst; ++e) {
            const LtOp& op = gops[g.rop_off + e];
            if (op.kind == KF_OP_MUL && (op.a < 0 || op.b < 0 || op.a >= g.nslots || op.b >= g.nslots || wlevel[op.a] == (1 << 30) || wlevel[op.b] == (1 << 30))) return -7;
            if (op.kind == KF_OP_VAR && (op.a < 0 || op.a >= g.nslots || wlevel[op.a] == (1 << 30))) return -8;
        }
        std::vector<double> sh(g.nslots);
        for (long long s = 0; s < rows; ++s) {
            for (int i = 0; i < nv; ++i) sh[i] = V[(size_t)i * rows + s];
            for (int e = 0; e < g.nops; ++e) {
                const LtOp& op = gops[g.op_off + e];
                KfOp o{};
                o.kind = op.kind; o.a = op.a; o.b = op.b; o.c = op.c;
                sh[op.j] = kf_eval_op(o, nv, prog.centres.data(), [&](int k) { return sh[k]; });
            }
            for (int e = 0; e < g.nst; ++e) {
                const LtOp& op = gops[g.rop_off + e];
                KfOp o{};
                o.kind = op.kind; o.a = op.a; o.b = op.b; o.c = op.c;
                out[(size_t)(g.row0 + e) * rows + s] = kf_eval_op(o, nv, prog.centres.data(), [&](int k) { return sh[k]; });
            }
        }
    }
    return next_row == nf ? 0 : -10;
}
