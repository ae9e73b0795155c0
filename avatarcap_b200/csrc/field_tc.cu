// (tcgen05 kernel -- under construction; see field_tc_stub.cu)
#include "common.cuh"
