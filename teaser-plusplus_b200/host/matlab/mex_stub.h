// Minimal declarations of the MATLAB MEX API used by teaser_mex.cc, for a compile (syntax/type) check in an image
// without MATLAB.  Never linked into anything.
#pragma once
#include <cstddef>
struct mxArray_tag;
typedef struct mxArray_tag mxArray;
enum mxComplexity { mxREAL = 0, mxCOMPLEX = 1 };
extern "C" {
bool mxIsDouble(const mxArray*);
bool mxIsComplex(const mxArray*);
bool mxIsLogicalScalar(const mxArray*);
bool mxIsLogicalScalarTrue(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
double* mxGetPr(const mxArray*);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateDoubleMatrix(size_t, size_t, mxComplexity);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexAtExit(void (*)(void));
}
