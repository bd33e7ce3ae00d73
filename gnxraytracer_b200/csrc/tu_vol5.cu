// Translation unit of libgnxrt.so: explicit instantiations of the GNX_TU_VOL5 kernels of gnx_kernels.cuh (see the
// "translation units" block at its end); compiled in parallel with the others by gnxraytracer_b200/build.py.
#define GNX_KERNELS_TEMPLATES_ONLY
#define GNX_TU_VOL5
#include "gnx_kernels.cuh"
