#pragma once
#include <vector>
#include "mex.h"
mxArray* t_double(mwSize m, mwSize n, const double* v);
mxArray* t_scalar(double v);
mxArray* t_logical(bool v);
mxArray* t_int32(mwSize m, mwSize n, const int32_t* v);
mxArray* t_uint16_3d(mwSize a0, mwSize a1, mwSize a2, const uint16_t* v);
mxArray* t_string(const char* s);
mxArray* t_sparse(mwSize m, mwSize n, const std::vector<mwIndex>& jc, const std::vector<mwIndex>& ir, const std::vector<double>& pr, bool logical);
mxArray* t_struct();
void t_setfield(mxArray* s, const char* name, mxArray* v);
void t_run_atexit();
extern int g_mex_locks;
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
