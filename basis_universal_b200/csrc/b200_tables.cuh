// b200_tables.cuh -- the device-resident copy of the codec tables (bu_tables.h), one per translation unit (no -rdc needed).
#pragma once
#include "bu_tables.h"
static __device__ const bu_tables d_tables =
#include "uastc_tables.inc"
;
