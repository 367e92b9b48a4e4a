// Functional stand-in for the parts of MATLAB's MEX API that matlab/cnmfe_b200_mex.cpp uses: mxArray is a malloc-backed
// column-major array (dense numeric, sparse double/logical, char, scalar struct).  mexErrMsgIdAndTxt throws.  Test only.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "mex.h"
#include "mex_test_api.h"

struct mxArray_tag {
    mxClassID cls = mxDOUBLE_CLASS;
    bool sparse = false;
    std::vector<mwSize> dims{0, 0};
    std::vector<unsigned char> data;
    std::vector<mwIndex> jc, ir;
    std::string str;
    std::vector<std::pair<std::string, mxArray*>> fields;
};
static size_t elsize(mxClassID c) {
    switch (c) {
        case mxDOUBLE_CLASS: case mxUINT64_CLASS: return 8;
        case mxINT32_CLASS: case mxSINGLE_CLASS: return 4;
        case mxUINT16_CLASS: return 2;
        default: return 1;
    }
}
static mwSize numel(const mxArray* a) { mwSize n = 1; for (mwSize d : a->dims) n *= d; return n; }
static void (*g_atexit)(void) = nullptr;
int g_mex_locks = 0;

extern "C" {
double* mxGetPr(const mxArray* a) { return a->data.empty() ? nullptr : (double*)a->data.data(); }
void* mxGetData(const mxArray* a) { return a->data.empty() ? nullptr : (void*)a->data.data(); }
mwIndex* mxGetJc(const mxArray* a) { return (mwIndex*)a->jc.data(); }
mwIndex* mxGetIr(const mxArray* a) { return (mwIndex*)a->ir.data(); }
mwSize mxGetM(const mxArray* a) { return a->dims[0]; }
mwSize mxGetN(const mxArray* a) { mwSize n = 1; for (size_t i = 1; i < a->dims.size(); ++i) n *= a->dims[i]; return n; }
double mxGetScalar(const mxArray* a) {
    if (a->data.empty()) throw std::runtime_error("mxGetScalar of an empty array");
    switch (a->cls) {
        case mxDOUBLE_CLASS: return *(const double*)a->data.data();
        case mxINT32_CLASS: return *(const int32_t*)a->data.data();
        case mxUINT64_CLASS: return (double)*(const uint64_t*)a->data.data();
        case mxUINT16_CLASS: return *(const uint16_t*)a->data.data();
        default: return *(const unsigned char*)a->data.data();
    }
}
bool mxIsEmpty(const mxArray* a) { return numel(a) == 0; }
bool mxIsSparse(const mxArray* a) { return a->sparse; }
bool mxIsLogical(const mxArray* a) { return a->cls == mxLOGICAL_CLASS; }
bool mxIsClass(const mxArray*, const char*) { return false; }
mxClassID mxGetClassID(const mxArray* a) { return a->cls; }
int mxGetString(const mxArray* a, char* buf, mwSize n) {
    if (a->cls != mxCHAR_CLASS || a->str.size() + 1 > n) return 1;
    std::strcpy(buf, a->str.c_str());
    return 0;
}
mxArray* mxGetField(const mxArray* s, mwIndex, const char* name) {
    if (s->cls != mxSTRUCT_CLASS) return nullptr;
    for (auto& f : s->fields) if (f.first == name) return f.second;
    return nullptr;
}
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID c, mxComplexity) {
    mxArray* a = new mxArray_tag();
    a->cls = c; a->dims = {m, n};
    a->data.assign(m * n * elsize(c), 0);
    return a;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, c); }
mxArray* mxDuplicateArray(const mxArray* a) { return new mxArray_tag(*a); }
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL); *mxGetPr(a) = v; return a; }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims.data(); }
mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->dims.size(); }
mwSize mxGetNumberOfElements(const mxArray* a) { return numel(a); }
bool mxIsStruct(const mxArray* a) { return a->cls == mxSTRUCT_CLASS; }
bool mxIsUint64(const mxArray* a) { return a->cls == mxUINT64_CLASS; }
bool mxIsInt32(const mxArray* a) { return a->cls == mxINT32_CLASS; }
bool mxIsSingle(const mxArray* a) { return a->cls == mxSINGLE_CLASS; }
bool mxIsChar(const mxArray* a) { return a->cls == mxCHAR_CLASS; }
bool mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
bool mxIsUint16(const mxArray* a) { return a->cls == mxUINT16_CLASS; }
bool mxIsUint8(const mxArray* a) { return a->cls == mxUINT8_CLASS; }
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char buf[2048];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    throw std::runtime_error(std::string(id) + ": " + buf);
}
void mexWarnMsgIdAndTxt(const char*, const char*, ...) {}
void mexLock(void) { ++g_mex_locks; }
void mexUnlock(void) { --g_mex_locks; }
int mexAtExit(void (*f)(void)) { g_atexit = f; return 0; }
}

// ---- construction helpers for the test driver
mxArray* t_double(mwSize m, mwSize n, const double* v) {
    mxArray* a = mxCreateDoubleMatrix(m, n, mxREAL);
    if (v && m * n) std::memcpy(a->data.data(), v, m * n * 8);
    return a;
}
mxArray* t_scalar(double v) { return mxCreateDoubleScalar(v); }
mxArray* t_logical(bool v) { mxArray* a = mxCreateNumericMatrix(1, 1, mxLOGICAL_CLASS, mxREAL); a->data[0] = v; return a; }
mxArray* t_int32(mwSize m, mwSize n, const int32_t* v) {
    mxArray* a = mxCreateNumericMatrix(m, n, mxINT32_CLASS, mxREAL);
    std::memcpy(a->data.data(), v, m * n * 4);
    return a;
}
mxArray* t_uint16_3d(mwSize a0, mwSize a1, mwSize a2, const uint16_t* v) {
    mxArray* a = mxCreateNumericMatrix(a0, a1 * a2, mxUINT16_CLASS, mxREAL);
    a->dims = {a0, a1, a2};
    std::memcpy(a->data.data(), v, a0 * a1 * a2 * 2);
    return a;
}
mxArray* t_string(const char* s) { mxArray* a = new mxArray_tag(); a->cls = mxCHAR_CLASS; a->str = s; a->dims = {1, std::strlen(s)}; return a; }
mxArray* t_sparse(mwSize m, mwSize n, const std::vector<mwIndex>& jc, const std::vector<mwIndex>& ir, const std::vector<double>& pr, bool logical) {
    mxArray* a = new mxArray_tag();
    a->cls = logical ? mxLOGICAL_CLASS : mxDOUBLE_CLASS; a->sparse = true; a->dims = {m, n};
    a->jc = jc; a->ir = ir;
    if (logical) a->data.assign(ir.size(), 1);
    else { a->data.resize(pr.size() * 8); std::memcpy(a->data.data(), pr.data(), pr.size() * 8); }
    return a;
}
mxArray* t_struct() { mxArray* a = new mxArray_tag(); a->cls = mxSTRUCT_CLASS; a->dims = {1, 1}; return a; }
void t_setfield(mxArray* s, const char* name, mxArray* v) { s->fields.emplace_back(name, v); }
void t_run_atexit() { if (g_atexit) g_atexit(); }
