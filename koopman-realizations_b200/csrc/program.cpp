// Host-side compiler: kf_basis -> KfProgram.  See program.h for the semantics.
#include "program.h"

#include <algorithm>

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

// Partition of the dictionary into dependency-closed feature groups of at most `max_slots` shared-memory slots
// (one group with slot = feature index if everything fits, or for dim_red, whose projection needs all features).
bool kf_build_lift_groups(const KfProgram& p, int max_slots, bool single, std::vector<LtOp>& gops, std::vector<LtStore>& gstore,
                          std::vector<LtGroup>& groups) {
    const int n = p.n_full(), nv = p.nv;
    std::vector<int> depth(n, 0);
    for (int j = 0; j < n; ++j) {
        const KfOp& op = p.ops[j];
        if (op.kind == KF_OP_MUL) {
            const int da = op.a < nv ? -1 : depth[op.a], db = op.b < nv ? -1 : depth[op.b];
            depth[j] = 1 + std::max(da, db);
        }
    }
    auto parents = [&](int j, int* out) -> int {       // features an op reads (variables excluded: always resident)
        const KfOp& op = p.ops[j];
        int c = 0;
        if (op.kind == KF_OP_MUL) { if (op.a >= nv) out[c++] = op.a; if (op.b >= nv) out[c++] = op.b; }
        return c;
    };
    gops.clear(); gstore.clear(); groups.clear();
    std::vector<char> in(n, 0);
    std::vector<int> members, stack, add, slot(n, -1);
    int j = 0;
    while (j < n) {
        // grow the group [j, j1) while its closure fits
        std::fill(in.begin(), in.end(), 0);
        members.clear();
        int count = nv, j1 = j;
        while (j1 < n) {
            stack.assign(1, j1);
            add.clear();
            while (!stack.empty()) {                   // closure of feature j1 not yet in the group
                const int f = stack.back(); stack.pop_back();
                if (f < nv || in[f]) continue;
                in[f] = 2; add.push_back(f);
                int pr[2];
                const int np = parents(f, pr);
                for (int q = 0; q < np; ++q) stack.push_back(pr[q]);
            }
            if (!single && j1 > j && count + (int)add.size() > max_slots) {
                for (int f : add) in[f] = 0;
                break;
            }
            for (int f : add) { in[f] = 1; members.push_back(f); }
            count += (int)add.size();
            ++j1;
        }
        // evaluation order: by depth, then by feature index; slots: variables 0 .. nv-1, then in that order (or identity)
        std::sort(members.begin(), members.end(), [&](int x, int y) { return depth[x] != depth[y] ? depth[x] < depth[y] : x < y; });
        // slots: variables 0 .. nv-1, then the parents that are not outputs of this group, then the outputs [max(j, nv), j1) in
        // feature order — the stored rows are ONE contiguous slot range (slot0 + e <-> row0 + e), so the store phase needs no table
        std::fill(slot.begin(), slot.end(), -1);
        for (int v = 0; v < nv; ++v) slot[v] = v;
        int next = nv;
        if (!single) {
            for (int f : members) if (f < j) slot[f] = next++;
            for (int f = std::max(j, nv); f < j1; ++f) slot[f] = next++;
        } else {
            for (int f : members) slot[f] = f;
        }
        LtGroup G{};
        G.op_off = (int)gops.size(); G.st_off = (int)gstore.size();
        G.nslots = single ? n : next;
        int lev = -1;
        for (int f : members) {
            while (lev < depth[f]) {
                if (G.nlevels >= KF_LT_MAXLEV) return false;
                G.level_start[G.nlevels++] = (int)gops.size() - G.op_off;
                ++lev;
            }
            const KfOp& op = p.ops[f];
            LtOp o{op.kind, op.a, op.b, slot[f], op.c};
            if (op.kind == KF_OP_MUL) { o.a = slot[op.a]; o.b = slot[op.b]; }
            gops.push_back(o);
        }
        G.level_start[G.nlevels] = (int)gops.size() - G.op_off;
        G.nops = (int)gops.size() - G.op_off;
        if (groups.empty())
            for (int v = 0; v < nv; ++v) gstore.push_back(LtStore{v, v});      // the variables are rows 0 .. nv-1 of psi
        for (int f = std::max(j, nv); f < j1; ++f) gstore.push_back(LtStore{slot[f], f});
        G.nst = (int)gstore.size() - G.st_off;
        G.slot0 = gstore[G.st_off].slot; G.row0 = gstore[G.st_off].row;
        groups.push_back(G);
        j = j1;
    }
    return true;
}


bool kf_build_lift_rowgroups(const KfProgram& p, int max_slots, int max_rows, std::vector<LtOp>& gops, std::vector<LtGroup>& groups) {
    const int n = p.n_full(), nv = p.nv;
    std::vector<int> depth(n, 0);
    for (int j = 0; j < n; ++j) {
        const KfOp& op = p.ops[j];
        if (op.kind == KF_OP_MUL) {
            const int da = op.a < nv ? -1 : depth[op.a], db = op.b < nv ? -1 : depth[op.b];
            depth[j] = 1 + std::max(da, db);
        }
    }
    gops.clear(); groups.clear();
    std::vector<char> in(n, 0), isp(n, 0);
    std::vector<int> members, parents, stack, add, newp, slot(n, -1);
    int j = 0;
    while (j < n) {
        for (int f : members) in[f] = 0;
        for (int f : parents) isp[f] = 0;
        members.clear(); parents.clear();
        int j1 = j;
        while (j1 < n && (max_rows <= 0 || j1 - j < max_rows)) {
            stack.assign(1, j1);
            add.clear(); newp.clear();
            while (!stack.empty()) {                   // closure of feature j1; every operand met on the way becomes a parent
                const int f = stack.back(); stack.pop_back();
                if (f < nv || in[f]) continue;
                in[f] = 1; add.push_back(f);
                const KfOp& op = p.ops[f];
                if (op.kind == KF_OP_MUL)
                    for (int q : {op.a, op.b})
                        if (q >= nv) {
                            if (!isp[q]) { isp[q] = 1; newp.push_back(q); }
                            stack.push_back(q);
                        }
            }
            if (j1 > j && nv + (int)parents.size() + (int)newp.size() > max_slots) {
                for (int f : add) in[f] = 0;
                for (int f : newp) isp[f] = 0;
                break;
            }
            members.insert(members.end(), add.begin(), add.end());
            parents.insert(parents.end(), newp.begin(), newp.end());
            ++j1;
        }
        std::sort(parents.begin(), parents.end(), [&](int x, int y) { return depth[x] != depth[y] ? depth[x] < depth[y] : x < y; });
        for (int v = 0; v < nv; ++v) slot[v] = v;
        int next = nv;
        for (int f : parents) slot[f] = next++;
        LtGroup G{};
        G.op_off = (int)gops.size();
        G.nslots = next;
        int lev = -1;
        for (int f : parents) {
            while (lev < depth[f]) {
                if (G.nlevels >= KF_LT_MAXLEV) return false;
                G.level_start[G.nlevels++] = (int)gops.size() - G.op_off;
                ++lev;
            }
            const KfOp& op = p.ops[f];
            LtOp o{op.kind, op.a, op.b, slot[f], op.c};
            if (op.kind == KF_OP_MUL) { o.a = slot[op.a]; o.b = slot[op.b]; }
            gops.push_back(o);
        }
        G.level_start[G.nlevels] = (int)gops.size() - G.op_off;
        G.nops = (int)gops.size() - G.op_off;
        G.rop_off = (int)gops.size();
        for (int f = j; f < j1; ++f) {
            const KfOp& op = p.ops[f];
            LtOp o{op.kind, op.a, op.b, -1, op.c};
            if (f < nv) o = LtOp{KF_OP_VAR, f, 0, -1, 0.0};
            else if (isp[f]) o = LtOp{KF_OP_VAR, slot[f], 0, -1, 0.0};
            else if (op.kind == KF_OP_MUL) { o.a = slot[op.a]; o.b = slot[op.b]; }
            gops.push_back(o);
        }
        G.st_off = 0; G.nst = j1 - j; G.row0 = j; G.slot0 = -1;
        groups.push_back(G);
        for (int f : parents) slot[f] = -1;
        j = j1;
    }
    return true;
}
