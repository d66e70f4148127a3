// Feature program: the observable dictionary psi = [v; blocks...; 1] compiled to a flat
// list of ops that a device thread evaluates in index order (one snapshot per thread).
//
// Semantics follow the reference's dictionary definitions (Ksysid.m:455-536, 629-863) and
// the `partitions` row order (partitions.m:206-219).  Floating-point association:
//   feature(row) = feature(row with its LAST non-zero entry zeroed) * primitive(LAST entry)
// and v_i^k = v_i^(k-1) * v_i, mirroring get_monomial's left-to-right product
// (Ksysid.m:687-690).  Products are plain IEEE multiplies (never fused).
#pragma once
#include <string>
#include <vector>
#include "../../include/koopfit.h"

enum : int { KF_OP_VAR = 0, KF_OP_CONST = 1, KF_OP_MUL = 2, KF_OP_COS = 3, KF_OP_SIN = 4, KF_OP_HERM = 5, KF_OP_GAUSS = 6 };

struct KfOp {
    int kind;   // KF_OP_*
    int a;      // VAR/COS/SIN/HERM: variable index; MUL: first factor; GAUSS: centre column
    int b;      // MUL: second factor; HERM: order
    int pad;
    double c;   // COS/SIN: angular multiplier 2*pi*j; CONST: value
};

struct KfProgram {
    int nv = 0;
    std::vector<KfOp> ops;          // n_full entries
    std::vector<double> centres;    // nv x ngauss, column-major (all gaussian blocks concatenated)
    int ngauss = 0;
    std::vector<double> pcs;        // n_full x n_pcs column-major (dim_red) or empty
    int n_pcs = 0;
    int n_full() const { return (int)ops.size(); }
    // lifted dimension seen by the fit: Ksysid.m:534, or nv + n_pcs + 1 after dim_red (1511-1517)
    int N() const { return n_pcs ? nv + n_pcs + 1 : n_full(); }
};

// ---- feature groups of the materialising lift (lift.cu): a contiguous range of output features plus everything they
// depend on (their dependency closure), renumbered into compact slots; slots 0 .. nv-1 are always the variables.
constexpr int KF_LT_MAXLEV = 32;
struct LtOp { int kind, a, b, j; double c; };      // op of a group in level order; a, b, j are SLOTS (GAUSS: a = centre column)
struct LtStore { int slot, row; };                 // a stored feature: slot in the group, output row = feature index
struct LtGroup {
    int op_off, nops;                  // ops of the group in the global op array
    int st_off, nst;                   // stored features in the global store array
    int nslots, nlevels;
    int slot0, row0;                   // the stored features are contiguous: slot0 + e holds output row row0 + e, e < nst
    int rop_off;                       // row groups (kf_build_lift_rowgroups): the nst row ops follow the slot ops in the op array
    int level_start[KF_LT_MAXLEV + 1]; // offsets into the group's ops; a level only reads slots written by earlier levels
};
// Partition into groups of at most max_slots slots (single: one group with slot = feature index, whatever its size).
// Returns false if a dependency chain is deeper than KF_LT_MAXLEV levels.
bool kf_build_lift_groups(const KfProgram& p, int max_slots, bool single, std::vector<LtOp>& gops, std::vector<LtStore>& gstore,
                          std::vector<LtGroup>& groups);

// Row groups of the streaming lift kernel: a group is a contiguous range of output rows [row0, row0 + nst); only the features
// that some member of the group READS (its parents) get a shared-memory slot and are evaluated level by level (ops
// [op_off, op_off + nops), j = slot); every output row then has ONE row op (gops[rop_off + e], operands = slots, j unused):
// a copy of a slot (KF_OP_VAR) or the feature's own op, evaluated in registers and stored straight to global memory — leaves
// (most monomials, every gaussian) never touch shared memory.  A group ends when its parents exceed max_slots or it has
// max_rows rows (max_rows <= 0: no limit).  Returns false if a dependency chain is deeper than KF_LT_MAXLEV levels.
bool kf_build_lift_rowgroups(const KfProgram& p, int max_slots, int max_rows, std::vector<LtOp>& gops, std::vector<LtGroup>& groups);

// rows of partitions(total, ones(1,nvars)) in the reference's order, appended to `out` (row-major)
void kf_partitions_ones(int total, int nvars, std::vector<int>& out);

// table of one block (rows x cols, row-major) as the reference enumerates it
int kf_block_rows(int type, int degree, int nv, int* rows, int* cols, std::vector<int>* table, std::string& err);

// compile a kf_basis; returns KF_OK or KF_EINVAL with `err` set
int kf_build_program(const kf_basis* basis, KfProgram& prog, std::string& err);

// regressor width (Ksysid.m:1019-1028); nw > 0: `loaded` model, the lifted state is [1; w] (x) psi (Ksysid.m:594-599)
inline int kf_regressor_width(int model, int N, int m, int nw = 0) {
    const int NL = N * (nw + 1);
    return model == KF_LINEAR ? NL + m : (model == KF_BILINEAR ? NL * (m + 1) : NL);
}
