/*
 * ddfa_b200.h — C ABI of libddfa_b200.so: the B200 (sm_100a) implementation of the DDFA
 * code_gnn GGNN hot path (embedding -> T x {edge gather-sum, GRU} -> attention readout -> MLP,
 * loss, backward, Adam).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is documented "host";
 *   - tensors are dense, row-major, fp32 activations/parameters, int32 graph structure;
 *   - `stream` is a cudaStream_t passed as void*; every call only ENQUEUES work on it
 *     (no allocation, no synchronisation, CUDA-graph-capture safe);
 *   - return value: 0 on success, negative ddfa_status otherwise; ddfa_last_error()
 *     returns a thread-local message for the last failing call;
 *   - the caller owns all memory; workspace sizes are reported by *_workspace_bytes().
 *
 * Notation: N nodes, E edges (DGL orientation: message src -> dst, aggregated at dst),
 * B graphs, K embedding tables (1 or 4), H embedding width, D = K*H hidden width,
 * T propagation steps, L MLP layers, V vocabulary size.
 *
 * Each entry cites the reference interface it replaces (paths relative to the DeepDFA repo).
 */
#ifndef DDFA_B200_H
#define DDFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDFA_ABI_VERSION 1

typedef enum ddfa_status {
  DDFA_OK = 0,
  DDFA_ERR_INVALID_ARG = -1,   /* bad pointer / size / unsupported shape            */
  DDFA_ERR_CUDA = -2,          /* a CUDA runtime call or launch failed              */
  DDFA_ERR_UNSUPPORTED = -3,   /* shape outside what the selected engine supports   */
  DDFA_ERR_WORKSPACE = -4      /* workspace too small                               */
} ddfa_status;

/* GEMM engines for the dense GRU matmuls */
#define DDFA_ENGINE_SIMT 0     /* fp32 FFMA reference kernels (any D % 4 == 0)              */
#define DDFA_ENGINE_TCGEN05 1  /* tcgen05 / TMEM, bf16x3 split operands, fp32 accumulate (D == 128) */

int ddfa_abi_version(void);
const char *ddfa_last_error(void);
/* 1 if the current device is compute capability 10.x, 0 otherwise, <0 on CUDA error */
int ddfa_device_supported(void);
/* 1 if the given DDFA_ENGINE_* is compiled into this library, else 0 */
int ddfa_engine_available(int engine);
/* Tuning knobs: process-wide selectors between equivalent launch configurations of the same kernels (defaults compiled in;
 * the library reads no environment variables).  ddfa_tuning_get returns -1 for an unknown key. */
enum {
  DDFA_TUNE_L2_HINTS = 0,       /* bit mask of L2 eviction-priority hints, default 23 (csrc/common.cuh) */
  DDFA_TUNE_PDL_MASK = 1,       /* bit mask of kernels launched with programmatic stream serialization, default 15 */
  DDFA_TUNE_GATHER_VARIANT = 2, /* launch shape of the D = 128 edge gather (ddfa_gather_sum_variant ids), default 9 */
  DDFA_TUNE_FWD_PAIR = 3,       /* 1: forward GRU kernel launched as 2-CTA clusters issuing tcgen05.mma.cta_group::2 (default 0) */
  DDFA_TUNE_GATE_BWD_TMA = 4,   /* gate backward: 0 register loads; 1 dh / gates / h stream through a TMA-fed shared-memory ring; 2 (default) = 1 + the folded gather's CSR scalars pipelined across iterations */
  DDFA_TUNE_GATHER_SRC_GROUPS = 5, /* image->image edge gather: row groups (of 4 rows) walked per warp with the CSR chain pipelined; 0 (default) = 1 group (2 / 4 measured neutral), or 1 / 2 / 4 */
  DDFA_TUNE__COUNT = 6
};
int ddfa_tuning_set(int key, int value);
int ddfa_tuning_get(int key);
/* development aid: in-kernel pipeline timeline of the tcgen05 kernels (SM-clock stamps per CTA / tile / event).
 * ddfa_debug_set(2, v): v = 0 off, 1 = forward + dgrad kernels, 2 = forward + wgrad kernels;
 * ddfa_debug_read(2 | 3, host, bytes): stamps of the backward (2) or forward (3) kernel's last launch;
 * ddfa_debug_read(4, host, 4): int32 count of timed-out mbarrier waits in the TMA-staged gather variants (0 when healthy). */
int ddfa_debug_set(int key, int value);
int ddfa_debug_read(int key, void *host_out, size_t bytes);
/* number of CUDA kernels this library has launched in this process (monotonic; for bench accounting) */
long long ddfa_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Graph structure.  Replaces the DGLGraph the reference hands to GatedGraphConv / pooling
 * (DDFA/code_gnn/models/flow_gnn/ggnn.py:95,102; batches built by dgl.batch,
 * DDFA/sastvd/linevd/dataset.py:76, datamodule.py:116-141).
 * ------------------------------------------------------------------------------------- */

/* COO -> CSR-by-destination (indptr/indices: in-neighbours of each node, sorted by source id)
 * and CSR-by-source (indptr_t/indices_t: out-neighbours, sorted; the transposed graph used by
 * the backward gather).  src/dst are int64 (idx_bytes=8, what DGL hands over) or int32
 * (idx_bytes=4).  indptr, indptr_t: int32[N+1]; indices, indices_t: int32[E].
 * Either output pair may be NULL to skip it.  Returns DDFA_ERR_INVALID_ARG for N<0/E<0. */
size_t ddfa_build_csr_workspace_bytes(int64_t num_edges, int32_t num_nodes);
int ddfa_build_csr(const void *src, const void *dst, int idx_bytes, int64_t num_edges,
                   int32_t num_nodes, int32_t *indptr, int32_t *indices, int32_t *indptr_t,
                   int32_t *indices_t, void *workspace, size_t workspace_bytes, void *stream);

/* batch_num_nodes int64[B] (DGLGraph.batch_num_nodes()) -> graph_ptr int32[B+1] (exclusive scan). */
int ddfa_graph_ptr(const int64_t *batch_num_nodes, int32_t num_graphs, int32_t *graph_ptr,
                   void *stream);

/* Batch producer: replaces the host-side collate — `dgl.batch([...])` in the GraphDataLoader
 * (DDFA/sastvd/linevd/datamodule.py:116-141) and `BigVulDatasetLineVD.get_indices` (DDFA/sastvd/linevd/dataset.py:63-76,
 * `dgl.batch([...]).to(device)`) — plus DGL's lazy CSR build, by slicing a device-resident ARENA of all graphs:
 *   arena = the CSR by destination and the CSR of the transposed graph over ALL graphs as one disjoint batch (what
 *   ddfa_build_csr produces for it), node_off int32[G+1] (first node of every graph), num_feats int64 feature vectors and
 *   the int32 _VULN vector over all nodes.
 * Out, for the graphs graph_ids[0..B) in that order: graph_ptr int32[B+1], indptr / indptr_t int32[N+1], indices /
 * indices_t int32[E], the feature vectors and _VULN restricted to the batch — bit-identical to ddfa_build_csr +
 * ddfa_graph_ptr on the collated COO of the same graphs.  batch_nodes / batch_edges = N and E of the batch (the caller
 * knows them from its host copy of the graph sizes; they size the outputs).  A bad id or inconsistent totals leave the
 * outputs untouched and raise the int32 counter at workspace[(B + 1) * 4].  feats / out_feats: host arrays of device
 * pointers, num_feats <= 8. */
size_t ddfa_arena_batch_workspace_bytes(int32_t batch_size);
int ddfa_arena_batch(const int32_t *graph_ids, int32_t batch_size, int32_t num_graphs, const int32_t *node_off,
                     const int32_t *indptr, const int32_t *indices, const int32_t *indptr_t, const int32_t *indices_t,
                     const int64_t *const *feats, int32_t num_feats, const int32_t *vuln, int32_t batch_nodes,
                     int32_t batch_edges, int32_t *out_graph_ptr, int32_t *out_indptr, int32_t *out_indices,
                     int32_t *out_indptr_t, int32_t *out_indices_t, int64_t *const *out_feats, int32_t *out_vuln,
                     void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * K1  embedding + concat.  Replaces ggnn.py:84-92 (4x nn.Embedding + torch.cat, or one).
 * idx[k]: int64[N] with values in [0,V); tables[k]: fp32[V,H]; x: fp32[N, K*H].
 * idx/tables are HOST arrays of K device pointers.  Out-of-range indices are clamped and
 * counted in *oob_count (int32 device counter, may be NULL) — the module raises on non-zero.
 * ------------------------------------------------------------------------------------- */
int ddfa_embed_concat_fwd(const int64_t *const *idx, const float *const *tables, int32_t num_tables,
                          int32_t vocab, int32_t width, int32_t num_nodes, float *x,
                          int32_t *oob_count, void *stream);
/* Same, and the rows also leave as h_0's activation image (row width K*H == 128; layout below, `image` holds
 * ddfa_act_image_bytes(num_nodes) bytes): what ddfa_act_to_image(x) would write, without the second pass over x.
 * Rows num_nodes .. (next multiple of 128) of the image are not written: the caller keeps them FINITE (e.g. zero-fills the
 * buffer once) — they are multiplied by the zero rows of the q images in the weight-gradient GEMM. */
int ddfa_embed_concat_fwd_image(const int64_t *const *idx, const float *const *tables, int32_t num_tables,
                                int32_t vocab, int32_t width, int32_t num_nodes, float *x, void *image,
                                int32_t *oob_count, void *stream);
/* dtables[k][idx_k[n], :] += (dx + dx2)[n, k*H:(k+1)*H]   (autograd of ggnn.py:84-92).
 * dx2 may be NULL; it lets the caller sum the two gradient paths into x (through the GGNN and
 * through the concat of ggnn.py:98) without a separate add kernel. */
int ddfa_embed_concat_bwd(const int64_t *const *idx, const float *dx, const float *dx2,
                          int32_t num_tables, int32_t vocab, int32_t width, int32_t num_nodes,
                          float *const *dtables, void *stream);

/* ---------------------------------------------------------------------------------------
 * K3  CSR edge gather-sum: out[v,:] = (accumulate ? out[v,:] : 0) + sum_{e in row v} h[indices[e],:]
 * Replaces DGL update_all(fn.copy_u('h','m'), fn.sum('m','a')) inside GatedGraphConv
 * (call site ggnn.py:95).  With the CSR-by-source arrays it is the backward of the same op.
 * D % 4 == 0, D <= 1024.  This is the HBM-roofline kernel (bytes: E*D*4 + N*D*4 + E*4 + (N+1)*4).
 * ------------------------------------------------------------------------------------- */
int ddfa_gather_sum(const int32_t *indptr, const int32_t *indices, const float *h,
                    int32_t num_nodes, int32_t dim, float *out, int accumulate, void *stream);
/* Tuning entry (scripts/gather_bench.py): same contract, explicit variant (D == 128): 0..9 register-path launch shapes,
 * 10 = neighbour rows staged in shared memory by per-row TMA bulk copies (cp.async.bulk + mbarrier), 11 = by tensor-map
 * tile::gather4 copies (four rows per UTMALDG) — csrc/gather_tma.cu. */
int ddfa_gather_sum_variant(int variant, const int32_t *indptr, const int32_t *indices, const float *h,
                            int32_t num_nodes, int32_t dim, float *out, int accumulate, void *stream);
/* The same gather for a source that exists only as its activation image (tcgen05 engine, steps t >= 1: h_t = hi + lo of the
 * image the forward GEMM read; no fp32 copy of h_t is kept).  D == 128. */
int ddfa_gather_sum_image_src(const int32_t *indptr, const int32_t *indices, const void *h_image,
                              int32_t num_nodes, int32_t dim, void *out_image, void *stream);


/* ---------------------------------------------------------------------------------------
 * Weight folding (done once per forward): w_fold = W_ih @ W  [3D,D], b_fold = W_ih @ b [3D]
 * so that  gi = (A h) w_fold^T + indeg * b_fold + b_ih  ==  GRUCell's  a W_ih^T + b_ih  with
 * a_v = sum_{u->v} (W h_u + b)   (DGL GatedGraphConv linears[0] + sum; ggnn.py:57-60).
 * ------------------------------------------------------------------------------------- */
int ddfa_fold_weights_fwd(const float *w_msg, const float *b_msg, const float *w_ih, int32_t dim,
                          float *w_fold, float *b_fold, void *stream);
/* dW_ih += dw_fold W^T + db_fold b^T ; dW += W_ih^T dw_fold ; db += W_ih^T db_fold */
int ddfa_fold_weights_bwd(const float *w_msg, const float *b_msg, const float *w_ih,
                          const float *dw_fold, const float *db_fold, int32_t dim, float *dw_msg,
                          float *db_msg, float *dw_ih, void *stream);

/* ---------------------------------------------------------------------------------------
 * K4  one GRU propagation step (torch.nn.GRUCell inside DGL GatedGraphConv; gate order r,z,n):
 *   gi = s w_fold^T + indeg b_fold + b_ih ; gh = h w_hh^T + b_hh
 *   r = sig(gi_r+gh_r) ; z = sig(gi_z+gh_z) ; n = tanh(gi_n + r*gh_n) ; h_out = (1-z)*n + z*h
 * s = gather-sum of h (K3).  indptr gives indeg.  If save_gates != NULL it receives
 * fp32[4][N][D] = r, z, n, gh_n (with b_hh_n) for the backward pass.
 * workspace: engine-dependent scratch (ddfa_gru_step_workspace_bytes).
 * ------------------------------------------------------------------------------------- */
size_t ddfa_gru_step_workspace_bytes(int32_t num_nodes, int32_t dim, int engine);
/* Once per forward (weights are constant over the T steps): engine-specific pre-packing of the
 * step's weights into `workspace` (tcgen05: bf16 hi/lo split, UMMA swizzled smem images; SIMT:
 * no-op).  The same workspace must then be passed to every ddfa_gru_step_fwd of that forward. */
int ddfa_gru_step_prepare(const float *w_fold, const float *b_fold, const float *b_ih,
                          const float *w_hh, const float *b_hh, int32_t dim, int engine,
                          void *workspace, size_t workspace_bytes, void *stream);
int ddfa_gru_step_fwd(const float *s, const float *h, const int32_t *indptr, const float *w_fold,
                      const float *b_fold, const float *b_ih, const float *w_hh, const float *b_hh,
                      int32_t num_nodes, int32_t dim, float *h_out, float *save_gates,
                      void *workspace, size_t workspace_bytes, int engine, void *stream);
/* ---- tcgen05 engine: activation images -------------------------------------------------------
 * With the tcgen05 engine (D == 128) activations travel between kernels as MMA-ready operands next to /
 * instead of fp32: image[node/128][hi|lo][cols 0-63 | 64-127] = 16 KB chunks of [128 rows x 64 bf16] in the
 * UMMA K-major SWIZZLE_128B layout, hi = bf16(x), lo = bf16(x - hi); rows past N are zero; the image has
 * exactly the size of the fp32 matrix rounded up to 128 rows (ddfa_act_image_bytes).  Producers:
 * ddfa_act_to_image (from fp32, used for h_0 = x), ddfa_gather_sum_image (s_t), ddfa_gru_step_fwd_image
 * (h_{t+1}).  Allocate images zero-initialised. */
size_t ddfa_act_image_bytes(int64_t num_nodes);
int ddfa_act_to_image(const float *x, int32_t num_nodes, int32_t dim, void *image, void *stream);
/* K3 writing the image of s (and, when out_f32 != NULL, also the fp32 matrix). */
int ddfa_gather_sum_image(const int32_t *indptr, const int32_t *indices, const float *h,
                          int32_t num_nodes, int32_t dim, void *out_image, float *out_f32, void *stream);
/* K4 on images: s_image / h_image in, h (fp32, for the z*h term) in; h_out fp32 and (optional) its image out.
 * workspace = the buffer prepared by ddfa_gru_step_prepare(engine = TCGEN05). */
int ddfa_gru_step_fwd_image(const void *s_image, const void *h_image, const float *h,
                            const int32_t *indptr, int32_t num_nodes, int32_t dim, float *h_out,
                            void *h_out_image, float *save_gates, const void *workspace,
                            size_t workspace_bytes, void *stream);
/* The form the training / inference drivers use from round 2 on (fewer bytes per step):
 *   h            fp32 [N,128] or NULL — NULL: the z*h term takes h from h_image (h = hi + lo, 2^-17 relative);
 *   h_out        fp32 or NULL (only the last step needs it, for the readout); h_out_image or NULL; at least one of the two;
 *   save_gates_packed  NULL, or ddfa_gru_gates_packed_bytes(N, D) bytes: per element one 64-bit word — r, z as 14-bit,
 *                n as 16-bit fixed point, gh_n as a 20-bit float (csrc/tc_common.cuh: pack_gates; <= 3.1e-5 error) — the four
 *                saved gate values as ONE 8-byte store instead of four fp32 planes. */
size_t ddfa_gru_gates_packed_bytes(int32_t num_nodes, int32_t dim);
int ddfa_gru_step_fwd_image_v2(const void *s_image, const void *h_image, const float *h,
                               const int32_t *indptr, int32_t num_nodes, int32_t dim, float *h_out,
                               void *h_out_image, void *save_gates_packed, const void *workspace,
                               size_t workspace_bytes, void *stream);

/* Backward of one step on images (tcgen05 engine): like ddfa_gru_step_bwd below, but s arrives as its
 * activation image (the one ddfa_gather_sum_image wrote in the forward pass); the q matrices and h are turned
 * into images inside the workspace.  workspace: ddfa_gru_step_bwd_workspace_bytes(N, D, TCGEN05), prepared by
 * ddfa_gru_step_prepare_bwd. */
/* h_image: the image of h (step input) kept from the forward pass, or NULL (it is then rebuilt in the workspace).
 * ds_prev / indptr_t / indices_t: NULL, or the incoming gradient is dh_out + A^T ds_prev — the transposed edge gather
 * (autograd of ggnn.py:95's message sum) of the ds the NEXT time step's call produced is folded into this call, with
 * A^T given as the CSR of the transposed graph (ddfa_build_csr).  ds must not alias ds_prev. */
int ddfa_gru_step_bwd_image(const float *dh_out, const float *ds_prev, const int32_t *indptr_t, const int32_t *indices_t,
                            const float *h, const void *h_image, const void *s_image,
                            const float *gates, const int32_t *indptr, int32_t num_nodes, int32_t dim, float *ds, float *dh,
                            float *dw_fold, float *db_fold, float *db_ih, float *dw_hh, float *db_hh,
                            void *workspace, size_t workspace_bytes, int wgrad_mode, void *stream);
/* Same with the saved state of ddfa_gru_step_fwd_image_v2: gates_packed instead of four fp32 planes; h may be NULL
 * (then h_image, required here, supplies h = hi + lo). */
int ddfa_gru_step_bwd_image_v2(const float *dh_out, const float *ds_prev, const int32_t *indptr_t, const int32_t *indices_t,
                               const float *h, const void *h_image, const void *s_image,
                               const void *gates_packed, const int32_t *indptr, int32_t num_nodes, int32_t dim, float *ds, float *dh,
                               float *dw_fold, float *db_fold, float *db_ih, float *dw_hh, float *db_hh,
                               void *workspace, size_t workspace_bytes, int wgrad_mode, void *stream);
/* wgrad_mode: 0 = dw_fold / dw_hh are updated before the call returns (stream order); 1 / 2 = deferred: the
 * weight-gradient GEMM accumulates per-CTA partial sums inside the workspace over the T steps of one backward pass
 * (1 = first step, overwrites; 2 = later steps) and ddfa_gru_step_bwd_finish adds them to dw_fold / dw_hh once. */
int ddfa_gru_step_bwd_finish(int32_t num_nodes, int32_t dim, float *dw_fold, float *dw_hh, void *workspace,
                             size_t workspace_bytes, void *stream);
/* wgrad_mode = DDFA_WGRAD_KEEP(slot): run no weight-gradient GEMM in the step call; its q images stay in slot `slot` of a
 * workspace sized by ddfa_gru_step_bwd_workspace_bytes_steps(N, D, TCGEN05, steps).  After the last step ONE call does the
 * weight-gradient GEMM of all kept steps (K = steps x nodes) and adds it to dw_fold / dw_hh:
 * s_images / h_images = host arrays of `steps` device pointers, entry t = the images of s_t and h_t that belong to slot t. */
#define DDFA_WGRAD_KEEP(slot) (16 + (slot))
#define DDFA_WGRAD_MAX_STEPS 16
size_t ddfa_gru_step_bwd_workspace_bytes_steps(int32_t num_nodes, int32_t dim, int engine, int32_t steps);
int ddfa_gru_bwd_wgrad_batched(const void *const *s_images, const void *const *h_images, int32_t steps, int32_t num_nodes,
                               int32_t dim, float *dw_fold, float *dw_hh, void *workspace, size_t workspace_bytes, void *stream);

/* Backward of one step.  In: dh_out, h (step input), s, gates.  Out: ds [N,D] (to be
 * transposed-gathered by the caller), dh [N,D] = dh_out*z + dgh W_hh (overwritten).
 * Accumulated (+=): dw_fold[3D,D], db_fold[3D], db_ih[3D], dw_hh[3D,D], db_hh[3D].
 * workspace: ddfa_gru_step_bwd_workspace_bytes(); with the tcgen05 engine it must first be
 * prepared once per backward pass by ddfa_gru_step_prepare_bwd (transposed bf16 hi/lo weight
 * images) and is then reused by every step of that pass. */
size_t ddfa_gru_step_bwd_workspace_bytes(int32_t num_nodes, int32_t dim, int engine);
int ddfa_gru_step_prepare_bwd(const float *w_fold, const float *w_hh, int32_t dim, int engine,
                              void *workspace, size_t workspace_bytes, void *stream);
int ddfa_gru_step_bwd(const float *dh_out, const float *h, const float *s, const float *gates,
                      const int32_t *indptr, const float *w_fold, const float *w_hh,
                      int32_t num_nodes, int32_t dim, float *ds, float *dh, float *dw_fold,
                      float *db_fold, float *db_ih, float *dw_hh, float *db_hh, void *workspace,
                      size_t workspace_bytes, int engine, void *stream);

/* ---------------------------------------------------------------------------------------
 * K3+K4 over all T steps: the whole dgl.nn.GatedGraphConv (ggnn.py:57-60 construction, :95 call) behind one call each.
 * They sequence the per-step entry points above (same kernels, same order as deepdfa_b200/engine.py) and carve every
 * intermediate out of ONE workspace of ddfa_ggnn_workspace_bytes(N, D, T, engine, training) bytes, which also carries
 * the saved activations from ddfa_ggnn_fwd(training = 1) to ddfa_ggnn_bwd (same workspace, untouched in between).
 *   fwd: x = h_0 [N,D] (the embedding output; must stay valid until the backward) -> h_out = h_T [N,D].
 *   bwd: dh_T [N,D] -> dx [N,D] = dL/dh_0 (overwritten); dw_msg[D,D], db_msg[D], dw_ih[3D,D], dw_hh[3D,D], db_ih[3D],
 *        db_hh[3D] accumulated (+=).  w_msg / b_msg = GatedGraphConv.linears[0], the rest = GatedGraphConv.gru.
 * engine = DDFA_ENGINE_SIMT (any D % 4 == 0) or DDFA_ENGINE_TCGEN05 (D == 128). */
size_t ddfa_ggnn_workspace_bytes(int32_t num_nodes, int32_t dim, int32_t n_steps, int engine, int training);
int ddfa_ggnn_fwd(const int32_t *indptr, const int32_t *indices, const float *x, int32_t num_nodes, int32_t dim,
                  int32_t n_steps, const float *w_msg, const float *b_msg, const float *w_ih, const float *w_hh,
                  const float *b_ih, const float *b_hh, float *h_out, void *workspace, size_t workspace_bytes,
                  int training, int engine, void *stream);
int ddfa_ggnn_bwd(const int32_t *indptr, const int32_t *indptr_t, const int32_t *indices_t, const float *x,
                  int32_t num_nodes, int32_t dim, int32_t n_steps, const float *w_msg, const float *b_msg,
                  const float *w_ih, const float *w_hh, const float *dh_T, float *dx, float *dw_msg, float *db_msg,
                  float *dw_ih, float *dw_hh, float *db_ih, float *db_hh, void *workspace, size_t workspace_bytes,
                  int engine, void *stream);

/* ---------------------------------------------------------------------------------------
 * K5-K7  readout + MLP.  Replaces torch.cat([ggnn_out, feat_embed]) (ggnn.py:98, never
 * materialised), DGL GlobalAttentionPooling(Linear(2D,1)) (ggnn.py:66-68,102) and the
 * output_layer MLP (ggnn.py:70-80,107).
 *   o_n = [h_T[n] | x[n]] ; g_n = <o_n, w_gate> + b_gate ; alpha = softmax of g over each graph
 *   pooled[b] = sum_n alpha_n o_n   (fp32[B,2D]; the encoder_mode output, ggnn.py:104-105)
 *   logits[b] = MLP(pooled[b])      (num_layers linears, ReLU between, last -> 1)
 * mlp_w / mlp_b: HOST arrays of num_layers device pointers ([2D,2D] ... [1,2D]); num_layers==0
 * skips the MLP (encoder mode; logits may be NULL).  Saved for backward when non-NULL:
 * gate_logit fp32[N], seg_max fp32[B], seg_sum fp32[B], mlp_act fp32[(L-1)][B][2D] (post-ReLU).
 * ------------------------------------------------------------------------------------- */
int ddfa_readout_mlp_fwd(const float *h_final, const float *x, const int32_t *graph_ptr,
                         int32_t num_graphs, int32_t dim, const float *w_gate, const float *b_gate,
                         const float *const *mlp_w, const float *const *mlp_b, int32_t num_layers,
                         float *pooled, float *logits, float *gate_logit, float *seg_max,
                         float *seg_sum, float *mlp_act, void *stream);
/* MLP backward: dlogits[B] -> dpooled[B,2D]; accumulates dmlp_w / dmlp_b (+=).
 * scratch: fp32[2][B][2D]. */
int ddfa_mlp_bwd(const float *dlogits, const float *pooled, const float *mlp_act,
                 const float *const *mlp_w, int32_t num_graphs, int32_t dim, int32_t num_layers,
                 float *dpooled, float *const *dmlp_w, float *const *dmlp_b, float *scratch,
                 void *stream);
/* Readout backward: dpooled[B,2D] -> dh_final[N,D], dx[N,D] (both overwritten);
 * accumulates dw_gate[2D], db_gate[1] (+=). */
int ddfa_readout_bwd(const float *dpooled, const float *pooled, const float *h_final, const float *x,
                     const int32_t *graph_ptr, int32_t num_graphs, int32_t dim, const float *w_gate,
                     const float *gate_logit, const float *seg_max, const float *seg_sum,
                     float *dh_final, float *dx, float *dw_gate, float *db_gate, void *stream);

/* ---------------------------------------------------------------------------------------
 * K8  graph labels + loss.  Replaces BaseModule.get_label (base_module.py:83-95: dgl.unbatch +
 * per-graph max of ndata["_VULN"]) and BCEWithLogitsLoss(pos_weight) (base_module.py:72-74,183).
 *   label[b] = max_n vuln[n] ; loss = (1/B) sum_b bce(logit_b, label_b; pos_weight)
 * dlogits[b] = grad_scale * d(sum_b bce)/dlogit_b  (caller passes grad_scale = 1/B_global).
 * loss_out: fp32[1] receives  loss_scale * sum_b bce  (caller passes loss_scale = 1/B_global).
 * vuln: int32[N].  labels: fp32[B] out.  dlogits may be NULL (evaluation).
 * ------------------------------------------------------------------------------------- */
int ddfa_graph_label_bce(const float *logits, const int32_t *vuln, const int32_t *graph_ptr,
                         int32_t num_graphs, float pos_weight, float loss_scale, float grad_scale,
                         float *labels, float *loss_out, float *dlogits, void *stream);
/* Same, for a batch padded to a bucket shape (FusedTrainer: one CUDA graph per bucket shape on a shuffled stream, the reference
 * reshuffles every epoch, datamodule.py:123-129): graphs [num_valid, num_graphs) are padding — they get a label but contribute
 * no loss term and dlogits = 0, so nothing of them reaches any gradient. */
int ddfa_graph_label_bce_valid(const float *logits, const int32_t *vuln, const int32_t *graph_ptr,
                               int32_t num_graphs, int32_t num_valid, float pos_weight, float loss_scale,
                               float grad_scale, float *labels, float *loss_out, float *dlogits, void *stream);

/* ---------------------------------------------------------------------------------------
 * K10  torch.optim.Adam(lr, betas, eps, weight_decay) with coupled L2 (DDFA/configs/
 * config_default.yaml:43-47) over one flat parameter buffer.  step_count: int32[1] device
 * counter, incremented by the kernel (graph-capture safe).
 * ------------------------------------------------------------------------------------- */
int ddfa_adam_flat(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                   int32_t *step_count, int64_t numel, float lr, float beta1, float beta2, float eps,
                   float weight_decay, void *stream);

/* ---------------------------------------------------------------------------------------
 * K10'  The data-parallel exchange fused with the optimizer over NVLink peer memory: ONE kernel per rank does
 * reduce-scatter (16-byte loads from every peer's gradient buffer) + Adam with coupled L2 on the rank's 1/R slice (moments are
 * sharded: exp_avg / exp_avg_sq are touched only inside the slice) + all-gather (16-byte stores of the new parameters into every
 * peer's parameter buffer), bracketed by two flag barriers in peer memory (csrc/allreduce_adam.cu).  Replaces
 * ncclAllReduce(flat gradients) + ddfa_adam_flat on every rank.  Collective: every rank launches it once per step.
 *   peer_params / peer_grads / peer_flags: HOST arrays of `world` device pointers, entry p = rank p's buffer as addressable from
 *   this device (symmetric allocation, peer-mapped); flags: >= 2 * world uint32 per rank, zero-initialised once;
 *   numel % 4 == 0; loss_offset: element index inside the gradient buffers of the per-rank loss word (summed into *loss_out,
 *   a LOCAL word; may be NULL); ticket: one zero-initialised local uint32; step_count as in ddfa_adam_flat (read as the
 *   barrier epoch, then incremented).  CUDA-graph capturable; waits are bounded (trap, not hang).
 * ------------------------------------------------------------------------------------- */
int ddfa_allreduce_adam_p2p(void *const *peer_params, const void *const *peer_grads, void *const *peer_flags,
                            int32_t rank, int32_t world, float *exp_avg, float *exp_avg_sq, int32_t *step_count,
                            int64_t numel, int64_t loss_offset, float *loss_out, uint32_t *ticket, float lr,
                            float beta1, float beta2, float eps, float weight_decay, void *stream);

/* ---------------------------------------------------------------------------------------
 * Generic row-major fp32 GEMM on the SIMT engine (building block, exported for tests):
 *   C[M,N] = alpha * op(A) op(B) + beta * C,  op(X) = X or X^T per trans flag.
 * split_k > 1 accumulates partial products with atomics (requires beta == 1, C pre-initialised).
 * ------------------------------------------------------------------------------------- */
int ddfa_sgemm(int trans_a, int trans_b, int32_t m, int32_t n, int32_t k, float alpha, const float *a,
               int32_t lda, const float *b, int32_t ldb, float beta, float *c, int32_t ldc,
               int32_t split_k, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DDFA_B200_H */
