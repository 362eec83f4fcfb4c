// Kernel flavour: double, energy, D4 (see d4b200_flavour.cuh).
#include "d4b200_handle.cuh"
#include "d4b200_small.cuh"
#define D4_TYPE double
#define D4_GRAD false
#define D4_S false
#include "d4b200_flavour.cuh"
D4_DEFINE_FLAVOUR(f64_e, D4_CLASSES_F64_E)
