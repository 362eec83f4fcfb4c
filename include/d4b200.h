/*
 * d4b200.h -- C ABI of the B200-native DFT-D4 hot path (libd4b200.so).
 *
 * Drop-in boundary: these entry points are what a maintainer of
 * the reference (tad-dftd4 v0.8.0, pure Python/PyTorch) would bind with
 * ``ctypes`` to replace, in one call, the chain
 *
 *   tad_mctc.ncoord.cn_d4                       (call site src/tad_dftd4/dispersion/base.py:390)
 *   D4Model.__init__/_get_alpha/trapzd          (src/tad_dftd4/model/base.py:107-151,367-431; utils.py:33-94)
 *   D4Model.weight_references (q and q=0)       (src/tad_dftd4/model/d4.py:103-228)
 *   D4Model.get_atomic_c6                       (src/tad_dftd4/model/d4.py:268-289)
 *   RationalDamping._f / dispersion2            (src/tad_dftd4/damping/functions.py:262-305; dispersion/twobody.py:89-201)
 *   ATM.calculate / get_atm_dispersion          (src/tad_dftd4/dispersion/threebody.py:54-163,210-256,305-321)
 *   torch.autograd over all of the above        (examples/forces.py:47-50)
 *
 * i.e. the body of ``Disp.calculate`` (dispersion/base.py:389-431) for the
 * default D4 method.  See INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *  - plain C, no torch types; every pointer named ``*_dev`` is DEVICE memory
 *    that the caller owns (borrowed for the duration of the call), pointers
 *    named ``*_host`` are host memory;
 *  - ``numbers`` is int64 [nbatch, nat] with 0 = padding, ``positions``
 *    [nbatch, nat, 3] in Bohr, ``q`` / ``energy`` [nbatch, nat]; real type is
 *    double (``_f64``) or float (``_f32``);
 *  - all device entry points are asynchronous on ``stream`` (a cudaStream_t
 *    passed as void*), allocate nothing and never synchronise;
 *  - return value: 0 = ok, < 0 = invalid argument (D4B200_E*), > 0 = CUDA
 *    runtime error code (cudaError_t).  Nothing throws.
 *  - per-structure problems found on the device (atomic number outside
 *    1..103, structure larger than the compiled paths support) are recorded
 *    in the workspace and read back with d4b200_status().
 */
#ifndef D4B200_H
#define D4B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D4B200_VERSION 100

/* invalid-argument codes */
#define D4B200_EINVAL (-1)      /* null pointer / negative size */
#define D4B200_EWORKSPACE (-2)  /* workspace too small */
#define D4B200_EPARAM (-3)      /* a1/a2 missing (NaN) etc. */
#define D4B200_ETABLE (-4)      /* table blob has the wrong size */
#define D4B200_EARCH (-5)       /* device is not sm_100 */
#define D4B200_ENUMBER (-6)     /* host-buffer entries: atomic number outside 1..103 */
#define D4B200_ETOOLARGE (-7)   /* host-buffer entries: structure beyond the small-family kernels */

/* device-side status bits returned by d4b200_status() */
#define D4B200_STATUS_BAD_NUMBER 1 /* atomic number outside 1..103 */
#define D4B200_STATUS_TOO_LARGE 2  /* structure too large for the available kernels */

/* model selector */
#define D4B200_MODEL_D4 0
#define D4B200_MODEL_D4S 1

/* number of size classes of the one-CTA-per-structure kernels */
#define D4B200_NCLASS 5

/* Flattened ``Param`` (src/tad_dftd4/damping/parameters/base.py:48-85) +
 * ``Cutoff`` (src/tad_dftd4/cutoff.py:36-90) + model weighting factor.
 * Defaults applied by the caller exactly as the reference does:
 * s6 = 1, s8 = 1 (dispersion/twobody.py:177-178), s9 = 1, alp = 16
 * (dispersion/threebody.py:233-234); ``has_s10`` mirrors `"s10" in param`
 * (dispersion/twobody.py:182). */
typedef struct d4b200_params {
  double s6, s8, s9, s10, a1, a2, alp;
  double disp2_cutoff; /* 60 */
  double disp3_cutoff; /* 40 */
  double cn_cutoff;    /* 30 (Cutoff.cn is NOT forwarded by the reference) */
  double wf;           /* Gaussian weighting factor, 6 (model/base.py:53) */
  int32_t has_s10;
  int32_t model; /* D4B200_MODEL_* */
} d4b200_params;

typedef struct d4b200_tables* d4b200_tables_t;

int d4b200_version(void);
const char* d4b200_error_string(int code);

/* Upload the per-element tables (tad_dftd4_b200/tables.py blob layout) to
 * ``device``; replaces the per-call model construction of
 * src/tad_dftd4/dispersion/base.py:363.  ``ga``/``gc`` are the charge-scaling
 * constants the blob was compiled with (model/base.py:51-52). */
int d4b200_tables_create(int device, const double* f64_blob_host, size_t n_f64,
                         const int32_t* i32_blob_host, size_t n_i32, double ga, double gc,
                         d4b200_tables_t* out);
int d4b200_tables_destroy(d4b200_tables_t tables);

/* Scratch the device entry points need for a [nbatch, nat] problem. */
size_t d4b200_workspace_bytes(int nbatch, int nat);

/* Atom-resolved dispersion energy, == tad_dftd4.dftd4(numbers, positions,
 * charge, param, q=q) (src/tad_dftd4/disp.py:44-146).  ``cn_dev`` (optional)
 * receives the coordination numbers. */
int d4b200_energy_f64(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                      const int64_t* numbers_dev, const double* positions_dev,
                      const double* q_dev, double* energy_dev, double* cn_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream);
int d4b200_energy_f32(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                      const int64_t* numbers_dev, const float* positions_dev, const float* q_dev,
                      float* energy_dev, float* cn_dev, void* workspace_dev,
                      size_t workspace_bytes, void* stream);

/* Host-buffer variant (the end-to-end call of bench.py): inputs and the result live in
 * HOST memory (pinned for full overlap).  The batch flows in ``chunks`` pieces (0 = auto)
 * through two internal pipeline slots, H2D copy -> kernels -> D2H copy, so the copies of
 * one chunk overlap the kernels of the other.  Synchronous. */
int d4b200_energy_host_f64(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                           const int64_t* numbers_host, const double* positions_host,
                           const double* q_host, double* energy_host, int chunks);
int d4b200_energy_host_f32(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                           const int64_t* numbers_host, const float* positions_host,
                           const float* q_host, float* energy_host, int chunks);
/* Host-buffer energy + forces: as d4b200_energy_host_* with the fused energy+gradient kernels per
 * chunk; ``grad_host`` [nbatch, nat, 3] receives d(sum_i E_i)/d positions (what
 * ``torch.autograd.grad(energy.sum(), positions)`` returns in examples/forces.py:47-50 of the
 * reference), ``gradq_host`` [nbatch, nat] (optional, may be NULL) d(sum_i E_i)/dq. */
int d4b200_energy_gradient_host_f64(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                                    const int64_t* numbers_host, const double* positions_host,
                                    const double* q_host, double* energy_host, double* grad_host,
                                    double* gradq_host, int chunks);
/* Host-buffer entry with NARROW atomic numbers: ``numbers_host`` holds ``numbers_itemsize``-byte integers
 * (8 = int64 as everywhere else, 4 = int32, 1 = uint8 -- every element fits one byte).  The array is
 * uploaded as it is (the caller keeps it narrow; there is no host-side conversion pass) and widened on
 * the device behind the copy: 1 instead of 8 of the 44 / 76 compulsory bytes per atom.
 * ``grad_host`` NULL = energy only, else the fused energy+gradient kernels as
 * d4b200_energy_gradient_host_*.  Like the other host entries the call fails with D4B200_ENUMBER /
 * D4B200_ETOOLARGE when a kernel flagged its input; ``status_out`` (optional) receives the raw
 * D4B200_STATUS_* bits.  Replaces: the padded ``numbers`` tensor of ``tad_dftd4.dftd4``
 * (disp.py:44-146 of the reference) when it lives on the host. */
int d4b200_energy_host_z_f64(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                             const void* numbers_host, int numbers_itemsize, const double* positions_host,
                             const double* q_host, double* energy_host, double* grad_host,
                             double* gradq_host, int chunks, int* status_out);
int d4b200_energy_host_z_f32(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                             const void* numbers_host, int numbers_itemsize, const float* positions_host,
                             const float* q_host, float* energy_host, float* grad_host, float* gradq_host,
                             int chunks, int* status_out);
int d4b200_energy_gradient_host_f32(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                                    const int64_t* numbers_host, const float* positions_host,
                                    const float* q_host, float* energy_host, float* grad_host,
                                    float* gradq_host, int chunks);

/* Vector-Jacobian product of the energy: for upstream weights
 * g = dL/dE [nbatch, nat] (NULL = all ones, i.e. L = sum E) returns
 * dL/dpositions [nbatch, nat, 3] and dL/dq [nbatch, nat] (either may be
 * NULL).  Replaces torch.autograd over the reference's dense tape
 * (examples/forces.py:47-50). */
int d4b200_gradient_f64(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                        const int64_t* numbers_dev, const double* positions_dev,
                        const double* q_dev, const double* grad_energy_dev,
                        double* grad_positions_dev, double* grad_q_dev, void* workspace_dev,
                        size_t workspace_bytes, void* stream);
int d4b200_gradient_f32(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                        const int64_t* numbers_dev, const float* positions_dev,
                        const float* q_dev, const float* grad_energy_dev,
                        float* grad_positions_dev, float* grad_q_dev, void* workspace_dev,
                        size_t workspace_bytes, void* stream);

/* Fused call: atom-resolved energies AND the vector-Jacobian product for the upstream
 * weights ``grad_energy_dev`` (NULL = ones, i.e. the forces of sum E) in one launch per
 * size class -- what `E = dftd4(...); torch.autograd.grad(E.sum(), positions)` needs. */
int d4b200_energy_gradient_f64(d4b200_tables_t tables, const d4b200_params* par, int nbatch,
                               int nat, const int64_t* numbers_dev, const double* positions_dev,
                               const double* q_dev, const double* grad_energy_dev,
                               double* energy_dev, double* grad_positions_dev, double* grad_q_dev,
                               void* workspace_dev, size_t workspace_bytes, void* stream);
int d4b200_energy_gradient_f32(d4b200_tables_t tables, const d4b200_params* par, int nbatch,
                               int nat, const int64_t* numbers_dev, const float* positions_dev,
                               const float* q_dev, const float* grad_energy_dev, float* energy_dev,
                               float* grad_positions_dev, float* grad_q_dev, void* workspace_dev,
                               size_t workspace_bytes, void* stream);

/* Model properties, == tad_dftd4.get_properties (src/tad_dftd4/disp.py:149-197) with
 * explicit charges: coordination numbers [nbatch, nat], pair C6 [nbatch, nat, nat]
 * (D4Model.get_atomic_c6 with the q-dependent weights) and static polarizabilities
 * [nbatch, nat] (BaseModel.get_polarizabilities, model/base.py:286-302).
 * ``energy_scratch_dev`` is a [nbatch, nat] scratch the kernel zeroes. */
int d4b200_properties_f64(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                          const int64_t* numbers_dev, const double* positions_dev,
                          const double* q_dev, double* cn_dev, double* c6_dev, double* alpha_dev,
                          double* energy_scratch_dev, void* workspace_dev, size_t workspace_bytes,
                          void* stream);
int d4b200_properties_f32(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                          const int64_t* numbers_dev, const float* positions_dev, const float* q_dev,
                          float* cn_dev, float* c6_dev, float* alpha_dev, float* energy_scratch_dev,
                          void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- single large structures (tiled kernels, no N^2 / N^3 storage) ---------
 * One structure of ``nat`` atoms (no padding rows; 0 entries are skipped), atoms
 * preferably in a spatially coherent order.  The call ACCUMULATES into
 * ``energy_dev`` [nat] (zero it first) the contributions of
 *   - two-body rows  [row_begin, row_end)            (E_i of these atoms, complete)
 *   - ATM centre groups [group_begin, group_end)     (groups of d4b200_large_group_size()
 *     consecutive atoms; shares of ALL atoms i,k that have a centre in these groups)
 * so that ranks owning disjoint ranges obtain the reference's
 * dftd4(numbers, positions, ...) by one all-reduce(sum) of ``energy_dev``
 * (SURVEY.md 8e).  ``cn_dev`` (optional) receives the coordination numbers,
 * ``group_cost_dev`` (optional, [ceil(nat/group)]) the neighbour count of every centre
 * group (cost of a group ~ count^2) for load-balanced partitioning; with
 * ``energy_dev == NULL`` only these are computed. */
int d4b200_large_group_size(void);
size_t d4b200_large_workspace_bytes(d4b200_tables_t tables, int nat, int fp32);
int d4b200_large_energy_f64(d4b200_tables_t tables, const d4b200_params* par, int nat,
                            const int64_t* numbers_dev, const double* positions_dev,
                            const double* q_dev, int row_begin, int row_end, int group_begin,
                            int group_end, double* energy_dev, double* cn_dev,
                            int* group_cost_dev, void* workspace_dev, size_t workspace_bytes,
                            void* stream);
int d4b200_large_energy_f32(d4b200_tables_t tables, const d4b200_params* par, int nat,
                            const int64_t* numbers_dev, const float* positions_dev,
                            const float* q_dev, int row_begin, int row_end, int group_begin,
                            int group_end, float* energy_dev, float* cn_dev, int* group_cost_dev,
                            void* workspace_dev, size_t workspace_bytes, void* stream);

/* Gradient of a single large structure, L = sum_i g_i E_i, in two stages so that ranks
 * owning disjoint row / centre-group ranges need only two all-reduces:
 *   stage 1 (d4b200_large_gradient_*): ACCUMULATES into force_dev [nat,3], dcn_dev [nat],
 *           dq_dev [nat] (zero them first) the direct two-body and ATM terms of this
 *           rank's ranges;  -> all-reduce(dcn_dev), all-reduce(dq_dev)
 *   stage 2 (d4b200_large_cn_chain_*): adds the coordination-number chain rule for the
 *           rows of this rank, given the TOTAL dL/dcn;  -> all-reduce(force_dev).
 * dq_dev is dL/dq (returned to torch so that it can be chained through the charges). */
int d4b200_large_gradient_f64(d4b200_tables_t tables, const d4b200_params* par, int nat,
                              const int64_t* numbers_dev, const double* positions_dev,
                              const double* q_dev, const double* grad_energy_dev, int row_begin,
                              int row_end, int group_begin, int group_end, double* force_dev,
                              double* dcn_dev, double* dq_dev, void* workspace_dev,
                              size_t workspace_bytes, void* stream);
int d4b200_large_gradient_f32(d4b200_tables_t tables, const d4b200_params* par, int nat,
                              const int64_t* numbers_dev, const float* positions_dev,
                              const float* q_dev, const float* grad_energy_dev, int row_begin,
                              int row_end, int group_begin, int group_end, float* force_dev,
                              float* dcn_dev, float* dq_dev, void* workspace_dev,
                              size_t workspace_bytes, void* stream);
/* Stage 1 with the atomic energies from the same launches (fused energy + gradient call of the tiled
 * family; what ``energy = dftd4(...); autograd.grad(energy.sum(), positions)`` asks for,
 * examples/forces.py:47-50 of the reference): additionally ACCUMULATES into energy_dev [nat] the
 * two-body rows and ATM centre groups of this rank (zero it first; all-reduce over the ranks). */
int d4b200_large_energy_gradient_f64(d4b200_tables_t tables, const d4b200_params* par, int nat,
                                     const int64_t* numbers_dev, const double* positions_dev,
                                     const double* q_dev, const double* grad_energy_dev, int row_begin,
                                     int row_end, int group_begin, int group_end, double* energy_dev,
                                     double* force_dev, double* dcn_dev, double* dq_dev,
                                     void* workspace_dev, size_t workspace_bytes, void* stream);
int d4b200_large_energy_gradient_f32(d4b200_tables_t tables, const d4b200_params* par, int nat,
                                     const int64_t* numbers_dev, const float* positions_dev,
                                     const float* q_dev, const float* grad_energy_dev, int row_begin,
                                     int row_end, int group_begin, int group_end, float* energy_dev,
                                     float* force_dev, float* dcn_dev, float* dq_dev,
                                     void* workspace_dev, size_t workspace_bytes, void* stream);
int d4b200_large_cn_chain_f64(d4b200_tables_t tables, const d4b200_params* par, int nat,
                              const int64_t* numbers_dev, const double* positions_dev,
                              const double* dcn_total_dev, int row_begin, int row_end,
                              double* force_dev, void* stream);
int d4b200_large_cn_chain_f32(d4b200_tables_t tables, const d4b200_params* par, int nat,
                              const int64_t* numbers_dev, const float* positions_dev,
                              const float* dcn_total_dev, int row_begin, int row_end,
                              float* force_dev, void* stream);

/* ---- model-level entry points (the reference's D4Model / D4SModel methods) ---------------
 * weight_references: zeta(q) * Gaussian weights, == D4Model.weight_references(cn, q)
 * (src/tad_dftd4/model/d4.py:103-228) -> gw [nbatch, nat, 7], or for par->model ==
 * D4B200_MODEL_D4S == D4SModel.weight_references (model/d4s.py:109-266) -> gw
 * [nbatch, nat(m), nat(n), 7] (weights of atom n as seen by partner m).  ``cn_dev`` / ``q_dev``
 * may be NULL (= 0, like the reference's defaults); ``dgwdcn_dev`` / ``dgwdq_dev`` (optional,
 * same shape) receive the derivatives returned by with_dgwdcn / with_dgwdq.  Only
 * par->model and par->wf are read.  ``status_dev``: optional int for D4B200_STATUS_BAD_NUMBER. */
int d4b200_weight_references_f64(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                                 const int64_t* numbers_dev, const double* cn_dev, const double* q_dev,
                                 double* gw_dev, double* dgwdcn_dev, double* dgwdq_dev, int* status_dev,
                                 void* stream);
int d4b200_weight_references_f32(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                                 const int64_t* numbers_dev, const float* cn_dev, const float* q_dev,
                                 float* gw_dev, float* dgwdcn_dev, float* dgwdq_dev, int* status_dev,
                                 void* stream);
/* get_atomic_c6(gw) -> [nbatch, nat, nat]: einsum('ijab,ia,jb->ij', rc6, gw, gw) (model/d4.py:268-289),
 * or for D4S einsum('ijab,jia,ijb->ij', rc6, gw, gw) (model/d4s.py:268-290). */
int d4b200_atomic_c6_f64(d4b200_tables_t tables, int model, int nbatch, int nat, const int64_t* numbers_dev,
                         const double* gw_dev, double* c6_dev, void* stream);
int d4b200_atomic_c6_f32(d4b200_tables_t tables, int model, int nbatch, int nat, const int64_t* numbers_dev,
                         const float* gw_dev, float* c6_dev, void* stream);
/* get_weighted_pols(gw) -> [nbatch, nat, nfreq] with nfreq = 23 (model/d4.py:291-307);
 * nfreq = 1 gives get_polarizabilities(gw) -> [nbatch, nat] (model/base.py:286-302). */
int d4b200_weighted_pols_f64(d4b200_tables_t tables, int nbatch, int nat, int nfreq,
                             const int64_t* numbers_dev, const double* gw_dev, double* alpha_dev,
                             void* stream);
int d4b200_weighted_pols_f32(d4b200_tables_t tables, int nbatch, int nat, int nfreq,
                             const int64_t* numbers_dev, const float* gw_dev, float* alpha_dev,
                             void* stream);

/* ---- parameter gradients ------------------------------------------------
 * Vector-Jacobian product of the atom-resolved energy with respect to the damping
 * parameters: out[b] = d(sum_i g_bi E_bi)/d(s6, s8, s9, s10, a1, a2, alp), float64
 * [nbatch, 7], one row per structure.  Replaces torch autograd through
 * RationalDamping._f / dispersion2 / get_atm_dispersion with ``Param`` values that require
 * grad (test/test_grad/test_param.py:40-100 of the reference).  ``c6q`` / ``c60`` are the
 * pair C6 matrices [nbatch, nat, nat] built from weight_references(cn, q) and
 * weight_references(cn, None) (two-body and ATM flavours, dispersion/twobody.py:72-74,
 * dispersion/threebody.py:313-315) with d4b200_atomic_c6_*, so both models are served.
 * ``grad_energy`` is the upstream gradient [nbatch, nat] (NULL = ones).  The s10 entry is 0
 * unless ``has_s10`` is set (the reference has no tensor to differentiate otherwise). */
size_t d4b200_param_vjp_workspace_bytes(int nbatch, int nat);
int d4b200_param_vjp_f64(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                         const int64_t* numbers_dev, const double* positions_dev,
                         const double* c6q_dev, const double* c60_dev,
                         const double* grad_energy_dev, double* out_dev, void* workspace_dev,
                         size_t workspace_bytes, void* stream);
int d4b200_param_vjp_f32(d4b200_tables_t tables, const d4b200_params* par, int nbatch, int nat,
                         const int64_t* numbers_dev, const float* positions_dev,
                         const float* c6q_dev, const float* c60_dev,
                         const float* grad_energy_dev, double* out_dev, void* workspace_dev,
                         size_t workspace_bytes, void* stream);

/* ---- EEQ-2019 atomic partial charges (the step before the hot path) ---------------------
 * Replaces tad_multicharge.get_eeq_charges(numbers, positions, charge, cutoff=cutoff.cn_eeq)
 * (third-party tad-multicharge==0.5.0; call sites src/tad_dftd4/dispersion/base.py:401-407 and
 * src/tad_dftd4/disp.py:190) for padded batches of structures with nat <= d4b200_eeq_limit():
 * one CTA per structure builds the bordered Coulomb system in shared memory and solves it.
 * ``param_host`` = [5][87] doubles: chi, eta, kappa (CN scaling), charge width, covalent radius
 * (Bohr), index 0 = padding.  ``charge_dev`` [nbatch] total charges; ``q_dev`` [nbatch, nat]
 * (0 on padding).  ``status_dev`` (optional, one int the caller zeroes) receives
 * D4B200_STATUS_BAD_NUMBER when an atomic number is outside 0..86.  The arithmetic is float64
 * for both I/O types. */
typedef struct d4b200_eeq* d4b200_eeq_t;
int d4b200_eeq_create(int device, const double* param_host, size_t n, d4b200_eeq_t* out);
int d4b200_eeq_destroy(d4b200_eeq_t eeq);
int d4b200_eeq_limit(void);
long long d4b200_eeq_launch_count(void);
int d4b200_eeq_charges_f64(d4b200_eeq_t eeq, int nbatch, int nat, const int64_t* numbers_dev,
                           const double* positions_dev, const double* charge_dev, double cn_cutoff,
                           double* q_dev, int* status_dev, void* stream);
int d4b200_eeq_charges_f32(d4b200_eeq_t eeq, int nbatch, int nat, const int64_t* numbers_dev,
                           const float* positions_dev, const float* charge_dev, double cn_cutoff,
                           float* q_dev, int* status_dev, void* stream);
/* Vector-Jacobian product of the charges: grad_positions [nbatch, nat, 3] =
 * (dq/dpositions)^T grad_q for the charges ``q_dev`` returned by d4b200_eeq_charges_*
 * (what torch.autograd does on the reference's tape when q is not given). */
int d4b200_eeq_vjp_f64(d4b200_eeq_t eeq, int nbatch, int nat, const int64_t* numbers_dev,
                       const double* positions_dev, double cn_cutoff, const double* q_dev,
                       const double* grad_q_dev, double* grad_positions_dev, int* status_dev,
                       void* stream);
int d4b200_eeq_vjp_f32(d4b200_eeq_t eeq, int nbatch, int nat, const int64_t* numbers_dev,
                       const float* positions_dev, double cn_cutoff, const float* q_dev,
                       const float* grad_q_dev, float* grad_positions_dev, int* status_dev,
                       void* stream);

/* The same two calls with the factor of the bordered EEQ matrix kept between them: the charges call
 * leaves the eliminated upper triangle, the reciprocal pivots and the raw coordination numbers of every
 * structure in ``factor_dev`` (nbatch * d4b200_eeq_factor_doubles(nat) doubles, float64 for both I/O
 * types), and a VJP call for the SAME numbers/positions only substitutes its right-hand side -- the
 * backward pass of ``get_eeq_charges`` (tad_multicharge, call site
 * /root/reference/src/tad_dftd4/dispersion/base.py:401-407) then costs no second elimination.
 * ``factor_dev`` may be NULL (plain call). */
size_t d4b200_eeq_factor_doubles(int nat);
int d4b200_eeq_charges_factor_f64(d4b200_eeq_t eeq, int nbatch, int nat, const int64_t* numbers_dev,
                                  const double* positions_dev, const double* charge_dev,
                                  double cn_cutoff, double* q_dev, double* factor_dev,
                                  int* status_dev, void* stream);
int d4b200_eeq_charges_factor_f32(d4b200_eeq_t eeq, int nbatch, int nat, const int64_t* numbers_dev,
                                  const float* positions_dev, const float* charge_dev,
                                  double cn_cutoff, float* q_dev, double* factor_dev, int* status_dev,
                                  void* stream);
int d4b200_eeq_vjp_factor_f64(d4b200_eeq_t eeq, int nbatch, int nat, const int64_t* numbers_dev,
                              const double* positions_dev, double cn_cutoff, const double* q_dev,
                              const double* grad_q_dev, const double* factor_dev,
                              double* grad_positions_dev, int* status_dev, void* stream);
int d4b200_eeq_vjp_factor_f32(d4b200_eeq_t eeq, int nbatch, int nat, const int64_t* numbers_dev,
                              const float* positions_dev, double cn_cutoff, const float* q_dev,
                              const float* grad_q_dev, const double* factor_dev,
                              float* grad_positions_dev, int* status_dev, void* stream);

/* Synchronises ``stream`` and returns the device status bits recorded by the
 * last energy/gradient call that used ``workspace_dev``. */
int d4b200_status(void* workspace_dev, void* stream, int* status_bits_out);

/* Number of kernel launches the last energy/gradient call on this thread
 * issued (for bench.py's gpu_launches claim). */
int d4b200_last_launch_count(void);
/* Cumulative number of launches of all batched energy/gradient calls of this process. */
long long d4b200_total_launch_count(void);

/* --- measurement hooks used by bench.py (no effect on results) ----------- */
/* Record CUDA events around every hot-kernel launch of the following calls. */
int d4b200_profile_enable(d4b200_tables_t tables, int enable);
/* Duration (ms) of the last call's hot kernel per size class (-1 = not launched),
 * followed by the batch-preparation time and the whole call's device time. */
int d4b200_profile_read(d4b200_tables_t tables, float* ms_out /*[D4B200_NCLASS + 2]*/);
/* Inclusive atom-count bounds of the size classes for a kernel flavour. */
int d4b200_class_caps(d4b200_tables_t tables, int fp32, int grad, int* caps_out /*[D4B200_NCLASS]*/);
int d4b200_class_caps_model(d4b200_tables_t tables, int fp32, int grad, int model,
                            int* caps_out /*[D4B200_NCLASS]*/);
/* Largest structure (atoms) the one-CTA-per-structure kernels of a flavour accept; larger
 * structures go through the d4b200_large_* entry points.  model: D4B200_MODEL_D4 / _D4S. */
int d4b200_small_limit(d4b200_tables_t tables, int fp32, int grad, int model);
/* Measured FP64 FMA throughput (TFLOP/s) of the device: roofline denominator. */
int d4b200_measure_fp64_peak(d4b200_tables_t tables, void* scratch_dev, size_t scratch_bytes,
                             void* stream, double* tflops_out);

/* Development profiling: accumulate clock64() cycles per kernel phase (16 slots per
 * size class); ``out`` (optional) receives and resets the counters. */
int d4b200_phase_profile(d4b200_tables_t tables, int enable,
                         unsigned long long* out /*[D4B200_NCLASS * 16] or NULL*/);

#ifdef __cplusplus
}
#endif
#endif /* D4B200_H */
