/* rsb_cuda16.cu -- 16-lane build of the env kernels: two environments per warp (models with nv <= 16) */
#define RSB_LANES 16
#define RSB_TABLE_NAME rsb_table_16
#include "rsb_kernels.inl"
