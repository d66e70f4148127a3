/* TEST-ONLY stub of the subset of MATLAB's MEX C API that matlab/koopfit_mex.cpp uses (signatures as documented for
 * R2019a, non-interleaved complex API).  MATLAB is absent in the build image; this header lets the shim be COMPILED
 * (g++ -fsyntax-only) so that type errors and ABI drift against include/koopfit.h are caught.  Nothing links against it. */
#ifndef KOOPFIT_MEX_STUB_H
#define KOOPFIT_MEX_STUB_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS, mxDOUBLE_CLASS } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
double* mxGetPr(const mxArray*);
double mxGetScalar(const mxArray*);
bool mxIsDouble(const mxArray*);
bool mxIsComplex(const mxArray*);
bool mxIsChar(const mxArray*);
bool mxIsCell(const mxArray*);
bool mxIsStruct(const mxArray*);
bool mxIsEmpty(const mxArray*);
bool mxIsNumeric(const mxArray*);
bool mxIsLogicalScalarTrue(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxGetField(const mxArray*, mwIndex, const char*);
mxArray* mxGetCell(const mxArray*, mwIndex);
void mxSetCell(mxArray*, mwIndex, mxArray*);
void mxSetField(mxArray*, mwIndex, const char*, mxArray*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
void mxDestroyArray(mxArray*);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
mxArray* mxCreateCellMatrix(mwSize, mwSize);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
void mexLock(void);
int mexAtExit(void (*)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
