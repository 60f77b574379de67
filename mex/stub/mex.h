/* Minimal stand-in for MATLAB/Octave's mex.h, used ONLY to syntax-check the gateways in this image
 * (neither MATLAB nor Octave is installed here).  Real builds use the vendor header:
 *     mkoctfile --mex -I../include foo.cpp -L.. -lcalz        (Octave)
 *     mex -R2017b -I../include foo.cpp -L.. -lcalz            (MATLAB, separate-complex API) */
#ifndef CALZ_STUB_MEX_H
#define CALZ_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxUINT64_CLASS = 13 } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
double* mxGetPr(const mxArray*);
double* mxGetPi(const mxArray*);
mwIndex* mxGetJc(const mxArray*);
mwIndex* mxGetIr(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
double mxGetScalar(const mxArray*);
bool mxIsSparse(const mxArray*);
bool mxIsDouble(const mxArray*);
bool mxIsComplex(const mxArray*);
bool mxIsCell(const mxArray*);
bool mxIsEmpty(const mxArray*);
bool mxIsChar(const mxArray*);
bool mxIsLogicalScalarTrue(const mxArray*);
mxArray* mxGetCell(const mxArray*, mwIndex);
bool mxIsClass(const mxArray*, const char*);
mxArray* mxGetProperty(const mxArray*, mwIndex, const char*);
void* mxGetData(const mxArray*);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
int mexCallMATLAB(int, mxArray**, int, mxArray**, const char*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateCellMatrix(mwSize, mwSize);
void mxSetCell(mxArray*, mwIndex, mxArray*);
void mxDestroyArray(mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
void* mxMalloc(size_t);
void mxFree(void*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
void mexWarnMsgIdAndTxt(const char*, const char*, ...);
int mexPrintf(const char*, ...);
int mexAtExit(void (*)(void));
void mexLock(void);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
