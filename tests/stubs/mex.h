// Stand-in for MATLAB's mex.h so that tests can build matlab/cnmfe_b200_mex.cpp against include/cnmfe_b200.h without MATLAB:
// type-checks with -fsyntax-only, and -- linked with tests/stubs/mex_impl.cpp, where mxArray is a real malloc-backed array -- runs.
#pragma once
#include <cstddef>
#include <cstdint>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize; typedef size_t mwIndex;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
typedef enum { mxDOUBLE_CLASS, mxINT32_CLASS, mxUINT64_CLASS, mxUINT16_CLASS, mxUINT8_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxSTRUCT_CLASS, mxSINGLE_CLASS } mxClassID;
extern "C" {
double* mxGetPr(const mxArray*); void* mxGetData(const mxArray*); mwIndex* mxGetJc(const mxArray*); mwIndex* mxGetIr(const mxArray*);
mwSize mxGetM(const mxArray*); mwSize mxGetN(const mxArray*); double mxGetScalar(const mxArray*); bool mxIsEmpty(const mxArray*);
bool mxIsSparse(const mxArray*); bool mxIsLogical(const mxArray*); bool mxIsClass(const mxArray*, const char*); mxClassID mxGetClassID(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize); mxArray* mxGetField(const mxArray*, mwIndex, const char*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity); mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray* mxDuplicateArray(const mxArray*); mxArray* mxCreateDoubleScalar(double);
const mwSize* mxGetDimensions(const mxArray*); mwSize mxGetNumberOfDimensions(const mxArray*); mwSize mxGetNumberOfElements(const mxArray*);
bool mxIsStruct(const mxArray*); bool mxIsUint64(const mxArray*); bool mxIsInt32(const mxArray*); bool mxIsSingle(const mxArray*); bool mxIsChar(const mxArray*); bool mxIsDouble(const mxArray*); bool mxIsUint16(const mxArray*); bool mxIsUint8(const mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...); void mexLock(void); void mexUnlock(void); int mexAtExit(void (*)(void)); void mexWarnMsgIdAndTxt(const char*, const char*, ...);
}
