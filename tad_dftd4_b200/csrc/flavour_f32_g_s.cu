// Kernel flavour: float, gradient, D4S (see d4b200_flavour.cuh).
#include "d4b200_handle.cuh"
#include "d4b200_small.cuh"
#define D4_TYPE float
#define D4_GRAD true
#define D4_S true
#include "d4b200_flavour.cuh"
D4_DEFINE_FLAVOUR(f32_g_s, D4_CLASSES_F32_G_S)
