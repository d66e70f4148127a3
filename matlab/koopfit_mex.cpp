// koopfit_mex.cpp — thin MEX shim over libkoopfit.so (include/koopfit.h).  Marshalling only: MATLAB's
// column-major double buffers are passed straight to the C ABI.  Build on a MATLAB host with
//     mex -I../include koopfit_mex.cpp -L../koopman-realizations_b200/lib -lkoopfit
// UNTESTED HERE: the build image has neither MATLAB nor Octave (no mex.h); see INTEGRATION.md.
//
// [K, info, Px, Py] = koopfit_mex('fit', alpha, beta, u, model_type, desc, opts, want_regressors)
#include <string.h>

#include <vector>

#include "koopfit.h"
#include "mex.h"

static kf_ctx* g_ctx = NULL;
static void at_exit(void) {
    if (g_ctx) {
        kf_destroy(g_ctx);
        g_ctx = NULL;
    }
}

static int model_code(const mxArray* a) {
    char buf[32];
    mxGetString(a, buf, sizeof(buf));
    if (!strcmp(buf, "linear")) return KF_LINEAR;
    if (!strcmp(buf, "bilinear")) return KF_BILINEAR;
    if (!strcmp(buf, "nonlinear")) return KF_NONLINEAR;
    mexErrMsgIdAndTxt("koopfit:model", "Invalid model_type chosen. Must be linear, bilinear, or nonlinear.");
    return -1;
}

static int obs_code(const char* s) {
    if (!strcmp(s, "poly")) return KF_POLY;
    if (!strcmp(s, "fourier")) return KF_FOURIER;
    if (!strcmp(s, "fourier_sparser")) return KF_FOURIER_SPARSER;
    if (!strcmp(s, "gaussian")) return KF_GAUSSIAN;
    if (!strcmp(s, "hermite")) return KF_HERMITE;
    return -1;   // unknown types are ignored, as in Ksysid.m:486-501
}

static double field_scalar(const mxArray* s, const char* name, double dflt) {
    const mxArray* f = mxGetField(s, 0, name);
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (!g_ctx) {
        if (kf_create(&g_ctx, 0)) mexErrMsgIdAndTxt("koopfit:create", kf_last_error(NULL));
        mexLock();
        mexAtExit(at_exit);
    }
    if (nrhs < 8) mexErrMsgIdAndTxt("koopfit:args", "usage: koopfit_mex('fit', alpha, beta, u, model_type, desc, opts, want_reg)");
    kf_problem pr;
    pr.M = (long long)mxGetM(prhs[1]);
    pr.nzeta = (int)mxGetN(prhs[1]);
    pr.m = (int)mxGetN(prhs[3]);
    pr.model = model_code(prhs[4]);
    pr.alpha = mxGetPr(prhs[1]);
    pr.beta = mxGetPr(prhs[2]);
    pr.u = mxGetPr(prhs[3]);
    pr.pc_cols = 0;   /* full K, as the reference returns it */
    pr.reserved = 0;

    // dictionary descriptor
    const mxArray* desc = prhs[5];
    const mxArray* types = mxGetField(desc, 0, "types");
    const double* degs = mxGetPr(mxGetField(desc, 0, "degrees"));
    const mxArray* cen = mxGetField(desc, 0, "centres");
    const mxArray* pcs = mxGetField(desc, 0, "pcs");
    const int nv = pr.nzeta + (pr.model == KF_NONLINEAR ? pr.m : 0);
    std::vector<kf_block> blocks;
    size_t gauss_used = 0;
    for (mwSize i = 0; i < mxGetNumberOfElements(types); ++i) {
        char buf[32];
        mxGetString(mxGetCell(types, i), buf, sizeof(buf));
        const int code = obs_code(buf);
        if (code < 0) continue;
        kf_block b;
        b.type = code;
        b.degree = (int)degs[i];
        b.centres = NULL;
        if (code == KF_GAUSSIAN) {
            b.centres = mxGetPr(cen) + gauss_used * nv;   // zeta0 is nv x degree, column-major
            gauss_used += (size_t)b.degree;
        }
        blocks.push_back(b);
    }
    kf_basis bs;
    bs.nv = nv;
    bs.nblocks = (int)blocks.size();
    bs.blocks = blocks.data();
    bs.pcs = (pcs && !mxIsEmpty(pcs)) ? mxGetPr(pcs) : NULL;
    bs.n_pcs = bs.pcs ? (int)mxGetN(pcs) : 0;

    // solve options
    const mxArray* opts = prhs[6];
    kf_solve sv;
    memset(&sv, 0, sizeof(sv));
    sv.least_squares = (int)field_scalar(opts, "least_squares", 1);
    const mxArray* t = mxGetField(opts, 0, "t");
    sv.nt = t ? (int)mxGetNumberOfElements(t) : 0;
    sv.t = t ? mxGetPr(t) : NULL;
    sv.delay_constraint = (int)field_scalar(opts, "delay_constraint", 0);
    sv.n = (int)field_scalar(opts, "n", 0);
    sv.nd = (int)field_scalar(opts, "nd", 0);
    sv.psd_shift = (int)field_scalar(opts, "psd_shift", KF_PSD_AS_REFERENCE);
    const bool want_reg = mxIsLogicalScalarTrue(prhs[7]) || mxGetScalar(prhs[7]) != 0;

    int N = 0, P = 0;
    if (kf_basis_dims(&bs, pr.model, pr.m, NULL, &N, &P)) mexErrMsgIdAndTxt("koopfit:basis", kf_last_error(NULL));
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
    if (kf_fit(g_ctx, &bs, &pr, &sv, &out)) mexErrMsgIdAndTxt("koopfit:fit", kf_last_error(g_ctx));

    if (nlhs > 1) {
        const char* fn[] = {"rank", "ls_method_used", "psd_shift_applied", "min_pivot", "max_pivot", "t_lift_gram_ms", "t_solve_ms", "objective",
                            "qp_gap", "qp_capped"};
        plhs[1] = mxCreateStructMatrix(1, 1, 10, fn);
        mxSetField(plhs[1], 0, "rank", mxCreateDoubleScalar(out.info.rank));
        mxSetField(plhs[1], 0, "ls_method_used", mxCreateDoubleScalar(out.info.ls_method_used));
        mxSetField(plhs[1], 0, "psd_shift_applied", mxCreateDoubleScalar(out.info.psd_shift_applied));
        mxSetField(plhs[1], 0, "min_pivot", mxCreateDoubleScalar(out.info.min_pivot));
        mxSetField(plhs[1], 0, "max_pivot", mxCreateDoubleScalar(out.info.max_pivot));
        mxSetField(plhs[1], 0, "t_lift_gram_ms", mxCreateDoubleScalar(out.info.t_lift_gram_ms));
        mxSetField(plhs[1], 0, "t_solve_ms", mxCreateDoubleScalar(out.info.t_solve_ms));
        mxSetField(plhs[1], 0, "objective", obj);
        mxSetField(plhs[1], 0, "qp_gap", gap);
        mxSetField(plhs[1], 0, "qp_capped", mxCreateDoubleScalar(out.info.qp_capped));
    }
    if (nlhs > 2) plhs[2] = Px ? Px : mxCreateDoubleMatrix(0, 0, mxREAL);
    if (nlhs > 3) plhs[3] = Py ? Py : mxCreateDoubleMatrix(0, 0, mxREAL);
}
