/* translation unit of libct_oracle.so (TEST INFRASTRUCTURE ONLY) */
#include "ct_oracle.c"
#include "ct_oracle_qparams.c"
#include "ct_oracle_fp4.c"
#include "ct_oracle_convert.c"
