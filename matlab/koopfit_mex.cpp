// koopfit_mex.cpp — thin MEX shim over libkoopfit.so (include/koopfit.h).  Marshalling only: MATLAB's
// column-major double buffers are passed straight to the C ABI.  Build on a MATLAB host with
//     mex -I../include koopfit_mex.cpp -L../koopman-realizations_b200/lib -lkoopfit
// UNTESTED IN MATLAB: the build image has neither MATLAB nor Octave (no mex.h); tests/test_program.py compiles this file
// against a stub of the MEX API (tests/mex_stub/mex.h) so that at least the types and the ABI calls are checked.
//
// One multi-GPU handle (kf_create_multi over EVERY visible device) lives for the MATLAB session: 'fit' goes through
// kf_fit_multi, so a single MATLAB process reaches all GPUs (north_star: "host side stays MATLAB ... snapshots are sharded
// across the 8 GPUs with one NCCL allreduce"); the single-device entry points use device 0's context.
//
//   [K, info, Px, Py] = koopfit_mex('fit', alpha, beta, u, model_type, desc, opts, want_reg [, w])   get_Koopman (Ksysid.m:987-1092)
//   [K, info, scale]  = koopfit_mex('fit_series', t, y, u, nd, model_type, desc, opts)               + get_scale / get_zeta / pairs
//   Kcell             = koopfit_mex('fit_batch', problems)       problems: struct array (alpha, beta, u, model_type, desc, opts)
//   ysim              = koopfit_mex('rollout', desc, model_type, models, zeta0, u, nout)             val_* (Ksysid.m:1623-1879)
//   Psi               = koopfit_mex('lift', desc, nv, V)                                             lift.econ_full
//   [coeff, latent, mu] = koopfit_mex('pca', desc, nv, V)                                          pca(Psi) of dim_red (Ksysid.m:1498)
//   X                 = koopfit_mex('mldivide', A, B)                                                A \ B (Ksysid.m:1216)
//   B                 = koopfit_mex('mpc_costB', A, Bmodel, z, horizon)                              Kmpc.m:569-596
//   koopfit_mex('set_option', name, value);   n = koopfit_mex('devices')
#include <string.h>

#include <string>
#include <vector>

#include "koopfit.h"
#include "mex.h"

static kf_multi* g_mc = NULL;
static void at_exit(void) {
    if (g_mc) {
        kf_destroy_multi(g_mc);
        g_mc = NULL;
    }
}
static kf_ctx* ctx0(void) { return kf_multi_ctx(g_mc, 0); }

static void need(bool ok, const char* id, const char* msg) {
    if (!ok) mexErrMsgIdAndTxt(id, msg);
}
static const mxArray* dbl(const mxArray* a, const char* what) {   // a real double array
    if (!a || !mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("koopfit:args", "%s must be a real double array", what);
    return a;
}
static std::string str_of(const mxArray* a, const char* what) {
    char buf[64];
    if (!a || !mxIsChar(a) || mxGetString(a, buf, sizeof(buf))) mexErrMsgIdAndTxt("koopfit:args", "%s must be a string", what);
    return std::string(buf);
}
static int model_code(const mxArray* a) {
    const std::string s = str_of(a, "model_type");
    if (s == "linear") return KF_LINEAR;
    if (s == "bilinear") return KF_BILINEAR;
    if (s == "nonlinear") return KF_NONLINEAR;
    mexErrMsgIdAndTxt("koopfit:model", "Invalid model_type chosen. Must be linear, bilinear, or nonlinear.");   // Ksysid.m:103
    return -1;
}
static int obs_code(const std::string& s) {
    if (s == "poly") return KF_POLY;
    if (s == "fourier") return KF_FOURIER;
    if (s == "fourier_sparser") return KF_FOURIER_SPARSER;
    if (s == "gaussian") return KF_GAUSSIAN;
    if (s == "hermite") return KF_HERMITE;
    return -1;   // unknown types are ignored, as in Ksysid.m:486-501
}
static double field_scalar(const mxArray* s, const char* name, double dflt) {
    const mxArray* f = (s && mxIsStruct(s)) ? mxGetField(s, 0, name) : NULL;
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}

// dictionary descriptor struct {types (cellstr), degrees, centres (nv x total gaussian degree), pcs} -> kf_basis
struct Basis {
    std::vector<kf_block> blocks;
    kf_basis b;
};
static void parse_basis(const mxArray* desc, int nv, Basis& out) {
    need(desc && mxIsStruct(desc), "koopfit:desc", "desc must be a struct with fields types, degrees [, centres, pcs]");
    const mxArray* types = mxGetField(desc, 0, "types");
    const mxArray* degs = mxGetField(desc, 0, "degrees");
    const mxArray* cen = mxGetField(desc, 0, "centres");
    const mxArray* pcs = mxGetField(desc, 0, "pcs");
    need(types && mxIsCell(types), "koopfit:desc", "desc.types must be a cell array of strings (obs_type)");
    dbl(degs, "desc.degrees");
    need(mxGetNumberOfElements(degs) == mxGetNumberOfElements(types), "koopfit:desc", "inputs must be of the same size");   // Ksysid.m:465-467
    size_t gauss_total = 0;
    for (mwSize i = 0; i < mxGetNumberOfElements(types); ++i)
        if (obs_code(str_of(mxGetCell(types, i), "desc.types{i}")) == KF_GAUSSIAN) gauss_total += (size_t)mxGetPr(degs)[i];
    if (gauss_total) {
        dbl(cen, "desc.centres");
        need(mxGetNumberOfElements(cen) >= gauss_total * (size_t)nv, "koopfit:desc",
             "desc.centres must hold nv x (sum of the gaussian degrees) centres (append zeta0 of every def_gaussianLift call)");
    }
    size_t gauss_used = 0;
    for (mwSize i = 0; i < mxGetNumberOfElements(types); ++i) {
        const int code = obs_code(str_of(mxGetCell(types, i), "desc.types{i}"));
        if (code < 0) continue;
        kf_block blk;
        blk.type = code;
        blk.degree = (int)mxGetPr(degs)[i];
        blk.centres = NULL;
        if (code == KF_GAUSSIAN) {
            blk.centres = mxGetPr(cen) + gauss_used * (size_t)nv;   // zeta0 is nv x degree, column-major
            gauss_used += (size_t)blk.degree;
        }
        out.blocks.push_back(blk);
    }
    out.b.nv = nv;
    out.b.nblocks = (int)out.blocks.size();
    out.b.blocks = out.blocks.empty() ? NULL : &out.blocks[0];
    out.b.pcs = (pcs && !mxIsEmpty(pcs)) ? mxGetPr(dbl(pcs, "desc.pcs")) : NULL;
    out.b.n_pcs = out.b.pcs ? (int)mxGetN(pcs) : 0;
}
static void parse_solve(const mxArray* opts, kf_solve& sv) {
    memset(&sv, 0, sizeof(sv));
    need(opts && mxIsStruct(opts), "koopfit:opts", "opts must be a struct");
    sv.least_squares = (int)field_scalar(opts, "least_squares", 1);
    const mxArray* t = mxGetField(opts, 0, "t");
    sv.nt = (t && !mxIsEmpty(t)) ? (int)mxGetNumberOfElements(dbl(t, "opts.t")) : 0;
    sv.t = sv.nt ? mxGetPr(t) : NULL;
    need(sv.least_squares || sv.nt > 0, "koopfit:opts", "the QP branch needs opts.t (lasso * N, Ksysid.m:996)");
    sv.ls_method = (int)field_scalar(opts, "ls_method", KF_LS_AUTO);
    sv.pivot_tol = field_scalar(opts, "pivot_tol", 0.0);
    sv.delay_constraint = (int)field_scalar(opts, "delay_constraint", 0);
    sv.n = (int)field_scalar(opts, "n", 0);
    sv.nd = (int)field_scalar(opts, "nd", 0);
    sv.psd_shift = (int)field_scalar(opts, "psd_shift", KF_PSD_AS_REFERENCE);
}
static mxArray* info_struct(const kf_info& in, mxArray* obj, mxArray* gap) {
    const char* fn[] = {"rank", "ls_method_used", "passes", "psd_shift_applied", "min_pivot", "max_pivot", "cond_est", "refine_passes",
                        "t_lift_gram_ms", "t_solve_ms", "objective", "qp_gap", "qp_capped"};
    mxArray* s = mxCreateStructMatrix(1, 1, 13, fn);
    mxSetField(s, 0, "rank", mxCreateDoubleScalar(in.rank));
    mxSetField(s, 0, "ls_method_used", mxCreateDoubleScalar(in.ls_method_used));
    mxSetField(s, 0, "passes", mxCreateDoubleScalar(in.passes));
    mxSetField(s, 0, "psd_shift_applied", mxCreateDoubleScalar(in.psd_shift_applied));
    mxSetField(s, 0, "min_pivot", mxCreateDoubleScalar(in.min_pivot));
    mxSetField(s, 0, "max_pivot", mxCreateDoubleScalar(in.max_pivot));
    mxSetField(s, 0, "cond_est", mxCreateDoubleScalar(in.cond_est));
    mxSetField(s, 0, "refine_passes", mxCreateDoubleScalar(in.refine_passes));
    mxSetField(s, 0, "t_lift_gram_ms", mxCreateDoubleScalar(in.t_lift_gram_ms));
    mxSetField(s, 0, "t_solve_ms", mxCreateDoubleScalar(in.t_solve_ms));
    mxSetField(s, 0, "objective", obj);
    mxSetField(s, 0, "qp_gap", gap);
    mxSetField(s, 0, "qp_capped", mxCreateDoubleScalar(in.qp_capped));
    return s;
}
static int width_of(int model, int N, int m, int nw) {
    const int NL = N * (nw + 1);
    return model == KF_LINEAR ? NL + m : (model == KF_BILINEAR ? NL * (m + 1) : NL);
}

// ---------------------------------------------------------------------------------------------------------------
static void cmd_fit(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    need(nrhs >= 8, "koopfit:args", "usage: koopfit_mex('fit', alpha, beta, u, model_type, desc, opts, want_reg [, w])");
    const mxArray *al = dbl(prhs[1], "alpha"), *be = dbl(prhs[2], "beta"), *uu = dbl(prhs[3], "u");
    need(!mxIsEmpty(al), "koopfit:args", "alpha is empty");
    need(mxGetM(be) == mxGetM(al) && mxGetN(be) == mxGetN(al), "koopfit:args", "size(beta) must equal size(alpha)");
    need(mxGetN(uu) == 0 || mxGetM(uu) == mxGetM(al), "koopfit:args", "u must have one row per snapshot pair");
    kf_problem pr;
    memset(&pr, 0, sizeof(pr));
    pr.M = (long long)mxGetM(al);
    pr.nzeta = (int)mxGetN(al);
    pr.m = (int)mxGetN(uu);
    pr.model = model_code(prhs[4]);
    pr.alpha = mxGetPr(al);
    pr.beta = mxGetPr(be);
    pr.u = mxGetPr(uu);
    pr.pc_cols = 0;   /* full K, as the reference returns it */
    if (nrhs > 8 && !mxIsEmpty(prhs[8])) {   /* snapshotPairs.w of a loaded model (Ksysid.m:1006-1011) */
        const mxArray* w = dbl(prhs[8], "w");
        need(mxGetM(w) == mxGetM(al), "koopfit:args", "w must have one row per snapshot pair");
        pr.nw = (int)mxGetN(w);
        pr.w = mxGetPr(w);
    }
    Basis bs;
    parse_basis(prhs[5], pr.nzeta + (pr.model == KF_NONLINEAR ? pr.m : 0), bs);
    kf_solve sv;
    parse_solve(prhs[6], sv);
    const bool want_reg = mxIsLogicalScalarTrue(prhs[7]) || (mxIsNumeric(prhs[7]) && !mxIsEmpty(prhs[7]) && mxGetScalar(prhs[7]) != 0);
    int N = 0;
    if (kf_basis_dims(&bs.b, pr.model, pr.m, NULL, &N, NULL)) mexErrMsgIdAndTxt("koopfit:basis", "%s", kf_last_error(NULL));
    const int P = width_of(pr.model, N, pr.m, pr.nw);
    const mwSize nt = (mwSize)(sv.least_squares ? 1 : sv.nt);
    mwSize dims[3] = {(mwSize)P, (mwSize)P, nt};
    plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
    kf_result out;
    memset(&out, 0, sizeof(out));
    out.K = mxGetPr(plhs[0]);
    mxArray *Px = NULL, *Py = NULL;
    if (want_reg) {
        Px = mxCreateDoubleMatrix((mwSize)pr.M, (mwSize)P, mxREAL);
        Py = mxCreateDoubleMatrix((mwSize)pr.M, (mwSize)P, mxREAL);
        out.Px = mxGetPr(Px);
        out.Py = mxGetPr(Py);
    }
    mxArray* obj = mxCreateDoubleMatrix(nt, 1, mxREAL);
    mxArray* gap = mxCreateDoubleMatrix(nt, 1, mxREAL);   /* certified f(K) - f* per budget (kf_result.qp_gap) */
    out.objective = mxGetPr(obj);
    out.qp_gap = mxGetPr(gap);
    /* every visible GPU: shards + one NCCL all-reduce inside the library */
    if (kf_fit_multi(g_mc, &bs.b, &pr, &sv, &out)) mexErrMsgIdAndTxt("koopfit:fit", "%s", kf_multi_last_error(g_mc));
    if (nlhs > 1) plhs[1] = info_struct(out.info, obj, gap);
    if (nlhs > 2) plhs[2] = Px ? Px : mxCreateDoubleMatrix(0, 0, mxREAL);
    if (nlhs > 3) plhs[3] = Py ? Py : mxCreateDoubleMatrix(0, 0, mxREAL);
}

static void cmd_fit_series(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    need(nrhs >= 8, "koopfit:args", "usage: koopfit_mex('fit_series', t, y, u, nd, model_type, desc, opts)");
    const mxArray *t = dbl(prhs[1], "t"), *y = dbl(prhs[2], "y"), *u = dbl(prhs[3], "u");
    kf_series se;
    memset(&se, 0, sizeof(se));
    se.T = (long long)mxGetNumberOfElements(t);
    se.n = (int)mxGetN(y);
    se.m = (int)mxGetN(u);
    se.nd = (int)mxGetScalar(prhs[4]);
    se.model = model_code(prhs[5]);
    need((long long)mxGetM(y) == se.T && (se.m == 0 || (long long)mxGetM(u) == se.T), "koopfit:args", "y and u need one row per time stamp");
    se.t = mxGetPr(t);
    se.y = mxGetPr(y);
    se.u = mxGetPr(u);
    const int nzeta = se.n * (se.nd + 1) + se.m * se.nd;   // Ksysid.m:86
    Basis bs;
    parse_basis(prhs[6], nzeta + (se.model == KF_NONLINEAR ? se.m : 0), bs);
    kf_solve sv;
    parse_solve(prhs[7], sv);
    int N = 0, P = 0;
    if (kf_basis_dims(&bs.b, se.model, se.m, NULL, &N, &P)) mexErrMsgIdAndTxt("koopfit:basis", "%s", kf_last_error(NULL));
    const mwSize nt = (mwSize)(sv.least_squares ? 1 : sv.nt);
    mwSize dims[3] = {(mwSize)P, (mwSize)P, nt};
    plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
    kf_result out;
    memset(&out, 0, sizeof(out));
    out.K = mxGetPr(plhs[0]);
    mxArray* obj = mxCreateDoubleMatrix(nt, 1, mxREAL);
    mxArray* gap = mxCreateDoubleMatrix(nt, 1, mxREAL);
    out.objective = mxGetPr(obj);
    out.qp_gap = mxGetPr(gap);
    const char* sf[] = {"y_offset", "y_factor", "u_offset", "u_factor", "M"};
    mxArray* scl = mxCreateStructMatrix(1, 1, 5, sf);
    mxArray* yo = mxCreateDoubleMatrix(1, se.n, mxREAL);
    mxArray* yf = mxCreateDoubleMatrix(1, se.n, mxREAL);
    mxArray* uo = mxCreateDoubleMatrix(1, se.m, mxREAL);
    mxArray* uf = mxCreateDoubleMatrix(1, se.m, mxREAL);
    kf_scale sc;
    memset(&sc, 0, sizeof(sc));
    sc.y_offset = mxGetPr(yo); sc.y_factor = mxGetPr(yf); sc.u_offset = mxGetPr(uo); sc.u_factor = mxGetPr(uf);
    if (kf_fit_series(ctx0(), &bs.b, &se, &sv, &sc, &out)) mexErrMsgIdAndTxt("koopfit:fit_series", "%s", kf_last_error(ctx0()));
    mxSetField(scl, 0, "y_offset", yo); mxSetField(scl, 0, "y_factor", yf);
    mxSetField(scl, 0, "u_offset", uo); mxSetField(scl, 0, "u_factor", uf);
    mxSetField(scl, 0, "M", mxCreateDoubleScalar((double)sc.M));
    if (nlhs > 1) plhs[1] = info_struct(out.info, obj, gap);
    if (nlhs > 2) plhs[2] = scl;
}

// the evaluate_rand_models.m loop (45-144) in one call: problems(i) = struct(alpha, beta, u, model_type, desc, opts)
static void cmd_fit_batch(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    need(nrhs >= 2 && mxIsStruct(prhs[1]), "koopfit:args", "usage: Kcell = koopfit_mex('fit_batch', problems)");
    const mwSize np = mxGetNumberOfElements(prhs[1]);
    std::vector<Basis> bases(np);
    std::vector<const kf_basis*> bptr(np);
    std::vector<kf_problem> probs(np);
    std::vector<kf_solve> solves(np);
    std::vector<kf_result> outs(np);
    plhs[0] = mxCreateCellMatrix(np, 1);
    for (mwSize i = 0; i < np; ++i) {
        const mxArray *al = dbl(mxGetField(prhs[1], i, "alpha"), "problems.alpha"), *be = dbl(mxGetField(prhs[1], i, "beta"), "problems.beta"),
                      *uu = dbl(mxGetField(prhs[1], i, "u"), "problems.u");
        need(mxGetM(be) == mxGetM(al) && mxGetN(be) == mxGetN(al), "koopfit:args", "size(beta) must equal size(alpha)");
        kf_problem& pr = probs[i];
        memset(&pr, 0, sizeof(pr));
        pr.M = (long long)mxGetM(al); pr.nzeta = (int)mxGetN(al); pr.m = (int)mxGetN(uu);
        pr.model = model_code(mxGetField(prhs[1], i, "model_type"));
        pr.alpha = mxGetPr(al); pr.beta = mxGetPr(be); pr.u = mxGetPr(uu);
        parse_basis(mxGetField(prhs[1], i, "desc"), pr.nzeta + (pr.model == KF_NONLINEAR ? pr.m : 0), bases[i]);
        bptr[i] = &bases[i].b;
        parse_solve(mxGetField(prhs[1], i, "opts"), solves[i]);
        int P = 0;
        if (kf_basis_dims(bptr[i], pr.model, pr.m, NULL, NULL, &P)) mexErrMsgIdAndTxt("koopfit:basis", "%s", kf_last_error(NULL));
        mwSize dims[3] = {(mwSize)P, (mwSize)P, (mwSize)(solves[i].least_squares ? 1 : solves[i].nt)};
        mxArray* K = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        mxSetCell(plhs[0], i, K);
        memset(&outs[i], 0, sizeof(kf_result));
        outs[i].K = mxGetPr(K);
    }
    if (np && kf_fit_batch(ctx0(), (int)np, &bptr[0], &probs[0], &solves[0], &outs[0])) mexErrMsgIdAndTxt("koopfit:fit_batch", "%s", kf_last_error(ctx0()));
}

// models: struct array with A, B (linear, bilinear) or F (nonlinear), plus n, m, nzeta; zeta0, u: cell arrays per trial
static void cmd_rollout(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    need(nrhs >= 7, "koopfit:args", "usage: ysim = koopfit_mex('rollout', desc, model_type, models, zeta0, u, nout)");
    const int model = model_code(prhs[2]);
    const mxArray* ms = prhs[3];
    need(mxIsStruct(ms) && mxGetNumberOfElements(ms) >= 1, "koopfit:args", "models must be a non-empty struct array");
    need(mxIsCell(prhs[4]) && mxIsCell(prhs[5]) && mxGetNumberOfElements(prhs[4]) == mxGetNumberOfElements(prhs[5]), "koopfit:args",
         "zeta0 and u must be cell arrays with one entry per trial");
    const int nm = (int)mxGetNumberOfElements(ms), ntr = (int)mxGetNumberOfElements(prhs[4]);
    const int n = (int)field_scalar(ms, "n", 0), m = (int)field_scalar(ms, "m", 0), nzeta = (int)field_scalar(ms, "nzeta", 0);
    need(n > 0 && nzeta > 0, "koopfit:args", "models need the fields n, m, nzeta");
    Basis bs;
    parse_basis(prhs[1], nzeta + (model == KF_NONLINEAR ? m : 0), bs);
    int N = 0;
    if (kf_basis_dims(&bs.b, model, m, NULL, &N, NULL)) mexErrMsgIdAndTxt("koopfit:basis", "%s", kf_last_error(NULL));
    std::vector<kf_model> mdl(nm);
    for (int c = 0; c < nm; ++c) {
        memset(&mdl[c], 0, sizeof(kf_model));
        mdl[c].model = model; mdl[c].n = n; mdl[c].m = m; mdl[c].nzeta = nzeta; mdl[c].N = N;
        const mxArray *A = mxGetField(ms, c, "A"), *B = mxGetField(ms, c, "B"), *F = mxGetField(ms, c, "F");
        if (model == KF_NONLINEAR) {
            need(F && mxGetM(F) == (mwSize)nzeta && mxGetN(F) == (mwSize)N, "koopfit:args", "models.F must be nzeta x N");
            mdl[c].F = mxGetPr(dbl(F, "models.F"));
        } else {
            need(A && mxGetM(A) == (mwSize)N && mxGetN(A) == (mwSize)N, "koopfit:args", "models.A must be N x N");
            need(B && mxGetM(B) == (mwSize)N && mxGetN(B) == (mwSize)(model == KF_LINEAR ? m : N * m), "koopfit:args", "models.B has the wrong size");
            mdl[c].A = mxGetPr(dbl(A, "models.A"));
            mdl[c].B = mxGetPr(dbl(B, "models.B"));
        }
    }
    int nout = (int)mxGetScalar(prhs[6]);
    if (nout <= 0) nout = n;
    std::vector<int> T(ntr);
    std::vector<const double*> z0(ntr), up(ntr);
    std::vector<double*> yp((size_t)nm * ntr);
    plhs[0] = mxCreateCellMatrix(nm, ntr);
    for (int k = 0; k < ntr; ++k) {
        const mxArray *z = dbl(mxGetCell(prhs[4], k), "zeta0{k}"), *u = dbl(mxGetCell(prhs[5], k), "u{k}");
        need(mxGetNumberOfElements(z) == (mwSize)nzeta && mxGetN(u) == (mwSize)m, "koopfit:args", "zeta0{k} needs nzeta entries, u{k} m columns");
        T[k] = (int)mxGetM(u);
        z0[k] = mxGetPr(z);
        up[k] = mxGetPr(u);
        for (int c = 0; c < nm; ++c) {
            mxArray* y = mxCreateDoubleMatrix(T[k], nout, mxREAL);
            mxSetCell(plhs[0], c + (mwSize)k * nm, y);
            yp[(size_t)c * ntr + k] = mxGetPr(y);
        }
    }
    if (kf_rollout(ctx0(), &bs.b, nm, &mdl[0], ntr, &T[0], &z0[0], &up[0], nout, &yp[0])) mexErrMsgIdAndTxt("koopfit:rollout", "%s", kf_last_error(ctx0()));
}

static void cmd_lift(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    need(nrhs >= 4, "koopfit:args", "usage: Psi = koopfit_mex('lift', desc, nv, V)");
    const int nv = (int)mxGetScalar(prhs[2]);
    const mxArray* V = dbl(prhs[3], "V");
    need((int)mxGetN(V) == nv, "koopfit:args", "V must have nv columns");
    Basis bs;
    parse_basis(prhs[1], nv, bs);
    int N = 0;
    if (kf_basis_dims(&bs.b, KF_NONLINEAR, 0, NULL, &N, NULL)) mexErrMsgIdAndTxt("koopfit:basis", "%s", kf_last_error(NULL));
    plhs[0] = mxCreateDoubleMatrix(mxGetM(V), N, mxREAL);
    if (mxGetM(V) && kf_lift(ctx0(), &bs.b, (long long)mxGetM(V), mxGetPr(V), mxGetPr(plhs[0]))) mexErrMsgIdAndTxt("koopfit:lift", "%s", kf_last_error(ctx0()));
}

// [coeff, latent, mu] = koopfit_mex('pca', desc, nv, V): the pca call of get_econ_observables (Ksysid.m:1498) on the lifted rows of V
static void cmd_pca(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    need(nrhs >= 4, "koopfit:args", "usage: [coeff, latent, mu] = koopfit_mex('pca', desc, nv, V)");
    const int nv = (int)mxGetScalar(prhs[2]);
    const mxArray* V = dbl(prhs[3], "V");
    need((int)mxGetN(V) == nv && mxGetM(V) >= 2, "koopfit:args", "V must have nv columns and at least two rows");
    Basis bs;
    parse_basis(prhs[1], nv, bs);
    int nf = 0;
    if (kf_basis_dims(&bs.b, KF_NONLINEAR, 0, &nf, NULL, NULL)) mexErrMsgIdAndTxt("koopfit:basis", "%s", kf_last_error(NULL));
    plhs[0] = mxCreateDoubleMatrix(nf, nf, mxREAL);
    mxArray* latent = mxCreateDoubleMatrix(nf, 1, mxREAL);
    mxArray* mu = mxCreateDoubleMatrix(1, nf, mxREAL);
    if (kf_pca(ctx0(), &bs.b, (long long)mxGetM(V), mxGetPr(V), mxGetPr(mu), mxGetPr(latent), mxGetPr(plhs[0])))
        mexErrMsgIdAndTxt("koopfit:pca", "%s", kf_last_error(ctx0()));
    if (nlhs > 1) plhs[1] = latent; else mxDestroyArray(latent);
    if (nlhs > 2) plhs[2] = mu; else mxDestroyArray(mu);
}

static void cmd_mldivide(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    need(nrhs >= 3, "koopfit:args", "usage: [X, rank] = koopfit_mex('mldivide', A, B)");
    const mxArray *A = dbl(prhs[1], "A"), *B = dbl(prhs[2], "B");
    need(mxGetM(A) == mxGetM(B) && !mxIsEmpty(A) && !mxIsEmpty(B), "koopfit:args", "A and B need the same number of rows");
    plhs[0] = mxCreateDoubleMatrix(mxGetN(A), mxGetN(B), mxREAL);
    int rank = 0;
    if (kf_mldivide(ctx0(), (long long)mxGetM(A), (int)mxGetN(A), (int)mxGetN(B), mxGetPr(A), mxGetPr(B), mxGetPr(plhs[0]), NULL, &rank))
        mexErrMsgIdAndTxt("koopfit:mldivide", "%s", kf_last_error(ctx0()));
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(rank);
}

static void cmd_mpc_costB(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    need(nrhs >= 5, "koopfit:args", "usage: B = koopfit_mex('mpc_costB', A, Bmodel, z, horizon)");
    const mxArray *A = dbl(prhs[1], "A"), *Bm = dbl(prhs[2], "Bmodel"), *z = dbl(prhs[3], "z");
    const int N = (int)mxGetM(A), h = (int)mxGetScalar(prhs[4]);
    need(N > 0 && (int)mxGetN(A) == N && (int)mxGetM(Bm) == N && mxGetN(Bm) % N == 0, "koopfit:args", "A must be N x N and Bmodel N x (N m)");
    const int m = (int)(mxGetN(Bm) / N), nz = (int)mxGetM(z);
    need((int)mxGetN(z) == N && (nz == 1 || nz == h), "koopfit:args", "z must have N columns and 1 or `horizon` rows");
    plhs[0] = mxCreateDoubleMatrix((mwSize)N * (h + 1), (mwSize)m * h, mxREAL);
    if (kf_mpc_costB_bilinear(ctx0(), N, m, h, 1, mxGetPr(A), mxGetPr(Bm), nz, mxGetPr(z), mxGetPr(plhs[0])))
        mexErrMsgIdAndTxt("koopfit:mpc_costB", "%s", kf_last_error(ctx0()));
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (!g_mc) {
        if (kf_create_multi(&g_mc, NULL, 0)) mexErrMsgIdAndTxt("koopfit:create", "%s", kf_last_error(NULL));
        mexLock();
        mexAtExit(at_exit);
    }
    need(nrhs >= 1 && mxIsChar(prhs[0]), "koopfit:args", "first argument: 'fit' | 'fit_series' | 'fit_batch' | 'rollout' | 'lift' | 'pca' | 'mldivide' | 'mpc_costB' | 'set_option' | 'devices'");
    const std::string cmd = str_of(prhs[0], "command");
    if (cmd == "fit") cmd_fit(nlhs, plhs, nrhs, prhs);
    else if (cmd == "fit_series") cmd_fit_series(nlhs, plhs, nrhs, prhs);
    else if (cmd == "fit_batch") cmd_fit_batch(nlhs, plhs, nrhs, prhs);
    else if (cmd == "rollout") cmd_rollout(nlhs, plhs, nrhs, prhs);
    else if (cmd == "lift") cmd_lift(nlhs, plhs, nrhs, prhs);
    else if (cmd == "pca") cmd_pca(nlhs, plhs, nrhs, prhs);
    else if (cmd == "mldivide") cmd_mldivide(nlhs, plhs, nrhs, prhs);
    else if (cmd == "mpc_costB") cmd_mpc_costB(nlhs, plhs, nrhs, prhs);
    else if (cmd == "set_option") {
        need(nrhs >= 3, "koopfit:args", "usage: koopfit_mex('set_option', name, value)");
        if (kf_multi_set_option(g_mc, str_of(prhs[1], "name").c_str(), mxGetScalar(prhs[2]))) mexErrMsgIdAndTxt("koopfit:option", "unknown option");
    } else if (cmd == "devices") {
        plhs[0] = mxCreateDoubleScalar(kf_multi_size(g_mc));
    } else {
        mexErrMsgIdAndTxt("koopfit:args", "unknown command %s", cmd.c_str());
    }
}
