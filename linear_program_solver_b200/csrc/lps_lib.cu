// The one translation unit of liblps_b200.so: the kernels live in headers shared by both halves of the
// C-ABI, so the two API files are compiled together.
#include "lps_api.cu"
#include "lps_exact_api.cu"
#include "lps_batch_api.cu"
