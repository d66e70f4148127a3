// Host-side compiler: kf_basis -> KfProgram.  See program.h for the semantics.
#include "program.h"

#include <cmath>
#include <map>

namespace {

// partitions.m:206-219 with an all-ones candidate set: loop the LAST variable's count
// i = 0..total in the outermost loop and recurse on the remaining variables.
void partitions_rec(int total, int nvars, std::vector<int>& prefix_rev, int width, std::vector<int>& out) {
    // prefix_rev holds the counts already fixed for the trailing variables (last first)
    if (nvars == 1) {
        out.push_back(total);
        for (int k = (int)prefix_rev.size() - 1; k >= 0; --k) out.push_back(prefix_rev[k]);
        return;
    }
    if (total == 0) {
        for (int k = 0; k < nvars; ++k) out.push_back(0);
        for (int k = (int)prefix_rev.size() - 1; k >= 0; --k) out.push_back(prefix_rev[k]);
        return;
    }
    for (int i = 0; i <= total; ++i) {
        prefix_rev.push_back(i);
        partitions_rec(total - i, nvars - 1, prefix_rev, width, out);
        prefix_rev.pop_back();
    }
}

const double kTwoPi = 2.0 * 3.14159265358979323846;

// Append one feature per row with the "last non-zero entry" product rule.
// prim(pos, val) returns the op of the primitive for a row whose only non-zero entry is (pos,val).
template <class Prim>
void chain_block(KfProgram& prog, const std::vector<int>& rows, int nrows, int ncols,
                 std::map<std::vector<int>, int>& index, int skip_rows, Prim prim) {
    std::vector<int> key(ncols), head(ncols), tail(ncols);
    for (int r = skip_rows; r < nrows; ++r) {
        int last = -1, nnz = 0;
        for (int c = 0; c < ncols; ++c) {
            key[c] = rows[(size_t)r * ncols + c];
            if (key[c] != 0) { last = c; ++nnz; }
        }
        KfOp op{};
        if (nnz == 1) {
            op = prim(last, key[last]);
        } else {
            head = key; head[last] = 0;
            std::fill(tail.begin(), tail.end(), 0); tail[last] = key[last];
            op.kind = KF_OP_MUL; op.a = index.at(head); op.b = index.at(tail);
        }
        index[key] = (int)prog.ops.size();
        prog.ops.push_back(op);
    }
}

}  // namespace

void kf_partitions_ones(int total, int nvars, std::vector<int>& out) {
    std::vector<int> prefix;
    partitions_rec(total, nvars, prefix, nvars, out);
}

int kf_block_rows(int type, int degree, int nv, int* rows, int* cols, std::vector<int>* table, std::string& err) {
    if (nv <= 0 || degree < 0) { err = "kf_block_rows: nv must be > 0 and degree >= 0"; return KF_EINVAL; }
    std::vector<int> t;
    int ncols = nv;
    switch (type) {
        case KF_POLY:      // def_polyLift 645-648 (all degree blocks; the caller drops the first nv rows, 488)
        case KF_HERMITE:   // def_hermiteLift 848-851
            for (int k = 1; k <= degree; ++k) kf_partitions_ones(k, nv, t);
            break;
        case KF_FOURIER_SPARSER:   // def_fourierLift_sparser 747-751 (constant row removed)
            ncols = 2 * nv;
            for (int k = 1; k <= degree; ++k) kf_partitions_ones(k, 2 * nv, t);
            break;
        case KF_FOURIER: {  // def_fourierLift 708-724: digits base (1+2d), last variable fastest, index 0 removed
            const long long base = 1 + 2LL * degree;
            long long total = 1;
            for (int i = 0; i < nv; ++i) {
                total *= base;
                if (total > (1LL << 26)) { err = "fourier dictionary too large"; return KF_EINVAL; }
            }
            t.resize((size_t)(total - 1) * nv);
            for (long long idx = 1; idx < total; ++idx) {
                long long q = idx;
                for (int i = nv - 1; i >= 0; --i) { t[(size_t)(idx - 1) * nv + i] = (int)(q % base); q /= base; }
            }
            break;
        }
        case KF_GAUSSIAN:
            ncols = 0;
            break;
        default:
            err = "unknown observable type";
            return KF_EINVAL;
    }
    int nrows = (type == KF_GAUSSIAN) ? degree : (ncols ? (int)(t.size() / ncols) : 0);
    if (rows) *rows = nrows;
    if (cols) *cols = ncols;
    if (table) table->swap(t);
    return KF_OK;
}

int kf_build_program(const kf_basis* basis, KfProgram& prog, std::string& err) {
    if (!basis || basis->nv <= 0 || basis->nblocks < 0 || (basis->nblocks && !basis->blocks)) {
        err = "kf_basis: nv must be > 0 and blocks non-NULL";
        return KF_EINVAL;
    }
    const int nv = basis->nv;
    prog = KfProgram();
    prog.nv = nv;
    // first nzeta observables are the state itself (Ksysid.m:484)
    for (int i = 0; i < nv; ++i) prog.ops.push_back(KfOp{KF_OP_VAR, i, 0, 0, 0.0});

    for (int bi = 0; bi < basis->nblocks; ++bi) {
        const kf_block& blk = basis->blocks[bi];
        if (blk.degree < 0) { err = "obs_degree must be >= 0"; return KF_EINVAL; }
        std::vector<int> rows;
        int nrows = 0, ncols = 0;
        if (blk.type != KF_GAUSSIAN) {
            int rc = kf_block_rows(blk.type, blk.degree, nv, &nrows, &ncols, &rows, err);
            if (rc) return rc;
        }
        std::map<std::vector<int>, int> index;
        switch (blk.type) {
            case KF_POLY: {
                // degree-1 rows are v itself: seed them and skip (Ksysid.m:488)
                int skip = blk.degree >= 1 ? nv : 0;
                for (int i = 0; i < skip; ++i) {
                    std::vector<int> e(nv, 0);
                    e[i] = 1;
                    index[e] = i;
                }
                chain_block(prog, rows, nrows, ncols, index, skip, [&](int pos, int val) {
                    std::vector<int> lower(nv, 0);
                    lower[pos] = val - 1;
                    return KfOp{KF_OP_MUL, index.at(lower), pos, 0, 0.0};   // v^k = v^(k-1) * v
                });
                break;
            }
            case KF_HERMITE:
                chain_block(prog, rows, nrows, ncols, index, 0,
                            [&](int pos, int val) { return KfOp{KF_OP_HERM, pos, val, 0, 0.0}; });
                break;
            case KF_FOURIER:
                chain_block(prog, rows, nrows, ncols, index, 0, [&](int pos, int val) {
                    int j = (val + 1) / 2;     // poop(2j)=cos, poop(2j+1)=sin, 1-based (Ksysid.m:712-713)
                    return KfOp{(val % 2 == 1) ? KF_OP_COS : KF_OP_SIN, pos, 0, 0, kTwoPi * (double)j};
                });
                break;
            case KF_FOURIER_SPARSER:
                chain_block(prog, rows, nrows, ncols, index, 0, [&](int pos, int val) {
                    if (pos < nv) return KfOp{KF_OP_SIN, pos, 0, 0, kTwoPi * (double)val};   // 779
                    return KfOp{KF_OP_COS, pos - nv, 0, 0, kTwoPi * (double)val};             // 784
                });
                break;
            case KF_GAUSSIAN:
                if (blk.degree > 0 && !blk.centres) { err = "gaussian block needs centres (nv x degree)"; return KF_EINVAL; }
                for (int k = 0; k < blk.degree; ++k) {
                    prog.ops.push_back(KfOp{KF_OP_GAUSS, prog.ngauss + k, 0, 0, 0.0});
                    for (int i = 0; i < nv; ++i) prog.centres.push_back(blk.centres[(size_t)k * nv + i]);
                }
                prog.ngauss += blk.degree;
                break;
            default:
                err = "unknown observable type";
                return KF_EINVAL;
        }
    }
    // constant term at the end (Ksysid.m:505)
    prog.ops.push_back(KfOp{KF_OP_CONST, 0, 0, 0, 1.0});

    if (basis->pcs && basis->n_pcs > 0) {
        prog.n_pcs = basis->n_pcs;
        prog.pcs.assign(basis->pcs, basis->pcs + (size_t)prog.n_full() * basis->n_pcs);
    }
    return KF_OK;
}
