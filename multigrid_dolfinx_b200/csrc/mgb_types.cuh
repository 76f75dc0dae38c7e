// Internal types of the engine (one translation unit: mgb_engine.cu): device CSR operators, the per-level state,
// the peer-memory halo plan and the handle itself.  Nothing here is part of the C ABI.
#pragma once

namespace {

constexpr int THREADS = 256;
thread_local std::string g_create_error;

// Lossless dictionary coding of an operator's (column, value) stream (mgb_code.cuh).  The CSR arrays stay on the
// device (artefacts, fallback kernels); the coded copy is what the row-stream kernel moves through HBM.
struct Coded {
    int mode = 0;                    // 0 none, 1 pair codes: entry -> (col - row, value), 2 value codes (+ the int32 columns),
                                     // 3 row-pattern codes: row -> its whole list of (col - row, value),
                                     // 4 anchored row patterns: row -> (first column, its list of (col - first column, value))
    int32_t* anchor = nullptr;       // mode 4: device, i32[nrows + pad]: every row's first stored column
    unsigned char* codes = nullptr;  // u8[nnz + pad]: index into dict (mode 3: u8[nrows + pad]: index into phead)
    DictEnt* dict = nullptr;         // device, 256 entries (mode 3: npent pattern entries)
    int2* phead = nullptr;           // mode 3: device, 256 x {first entry, length}
    int npent = 0;                   // mode 3: entries in the pattern table (multiple of 8)
    DictEnt* dict_win = nullptr;     // mode 3, row-window kernel (k_rowwin): the pattern table again, every entry's offset expressed as
    WinPlan win{};                   // (window << WIN_GSHIFT) | (offset - window minimum); win.ng == 0: not available
    HotPlan hotplan{};               // mode 3: the most frequent pattern (handed to the row-window kernel as kernel parameters)
    uint32_t* pmask = nullptr;       // mode 3, hot-row kernel (k_hotrow): device, 256 x (bit mask over the hot pattern's entries | HOT_SLOW)
    HotArgs hot{};                   // ... and the hot pattern as kernel parameters; hot_ok: the kernel applies
    bool hot_ok = false;
    int hot_slow = 0;                // patterns that are not sub-patterns of the hot one (table walk)
    int ndict = 0;                   // entries (mode 3: patterns) in use
    int nvals = 0, ndeltas = 0;      // distinct values / distinct (col - row) found
};

struct DevCsr {
    int64_t nrows = 0, ncols = 0, nnz = 0;
    Coded cd;
    int ccfg = 0;            // row-stream kernel configuration (code_choice) when cd.mode != 0
    int wcfg = 0;            // > 0: row-window kernel configuration (win_choice): x staged in shared memory
    int hcfg = 0;            // > 0: hot-row kernel configuration (hot_choice): speculative loads at the hot pattern's offsets
    int32_t* rowptr = nullptr;
    int32_t* cols = nullptr;
    double* vals = nullptr;
    int32_t* tiles = nullptr;
    int ntiles = 0;
    int4* sdesc = nullptr;   // stream kernel: per-tile {row0, nrows, nz0a, nent}
    int sntiles = 0;
    int scfg = 0;            // stream kernel configuration chosen for THIS operator (from its row lengths)
    std::vector<int4> sdesc_host;
    // row-sharded operators: stream tiles split into [boundary-low | interior | boundary-high]; interior rows reference
    // no ghost column, so they can run while the halo exchange is in flight
    bool split = false;
    int64_t int_r0 = 0, int_r1 = 0;  // sharded: rows [int_r0, int_r1) reference no ghost column
    int t_int0 = 0, t_int1 = 0;      // interior tiles = sdesc[t_int0 .. t_int1)
    int4* sdesc_bnd = nullptr;       // boundary tiles, low block then high block
    int n_bnd = 0;
    int iter = 2;       // tile kernel: groups of 4 entries per thread
    int family = 1;     // 1 tile, 2 sub-warp
    int lpr = 4;        // sub-warp lanes per row
    int max_row = 0;
    std::vector<int32_t> break_tile;   // tile index at each row breakpoint (colour boundaries)
    bool present() const { return rowptr != nullptr; }
};

constexpr int P2P_MAX_PEERS = 16;
constexpr int P2P_FLAG_SLOTS = 256;          // one arrival flag per sender rank
struct P2PPlan {                             // passed to the halo kernels by value
    int npeers = 0;
    int send_off[P2P_MAX_PEERS + 1] = {0};   // prefix sums into send_idx
    int recv_off[P2P_MAX_PEERS + 1] = {0};   // prefix sums into the ghost section
    double* rstage[P2P_MAX_PEERS][2] = {};   // neighbour's staging copies, already offset to this rank's slot
    unsigned long long* rflag[P2P_MAX_PEERS] = {};   // neighbour's arrival flag for this rank
    int peer_rank[P2P_MAX_PEERS] = {0};
    unsigned long long* counters = nullptr;  // see Level::p2p_counters
    unsigned long long* flags = nullptr;     // this rank's arrival flags (indexed by sender rank)
    double* stage = nullptr;                 // this rank's staging copies (2 x n_ghost)
    int n_ghost = 0;
    // fused exchange (HaloFuse): the neighbours' iterate buffers and a second set of flags
    double* rvec[P2P_MAX_PEERS][2] = {};     // neighbour's v / vtmp, already offset to this rank's slot of its ghost section
    unsigned long long* rflag2[P2P_MAX_PEERS] = {};  // neighbour's fused-exchange flag for this rank
    unsigned long long* flags2 = nullptr;    // this rank's fused-exchange flags (indexed by sender rank)
};

struct P2PBlob {                             // what a rank publishes per level (mgb_p2p_export)
    cudaIpcMemHandle_t handle;
    long long n_ghost;
    long long n_owned;                       // the iterate buffers v / vtmp live in the arena too: byte offsets of their first entry
    long long off_vec[2];
    int npeers;
    int peer_rank[P2P_MAX_PEERS];
    int recv_off[P2P_MAX_PEERS + 1];
};

struct Level {
    int level = 0;
    int64_t n = 0;
    HostCsr A_host;                  // released after finalize
    HostCsr P_host, R_host;          // transfer from level-1 to this level (this level = fine side)
    std::vector<int32_t> inj_host;
    int r_mode = MGB_R_INJECTION;
    int dim_fw = 2;
    bool has_transfer = false;       // transfer pair (level-1, level) was set
    int64_t n_coarse = 0;            // coarse rows this rank produces when restricting from this level
    // ---- row-sharded (multi-GPU) state: this rank owns n rows; vectors hold n + n_ghost entries, ghosts last
    int64_t n_ghost = 0;
    bool stub = false;               // gathered level on a non-root rank: full-size vectors, no operators
    bool device_born = false;        // operators were generated on the device (mgb_synth_*): no host copy exists
    int syn_dim = 0, syn_m = 0;      // geometry of a generated level
    int64_t row_begin = 0, row_end = 0, ghost_lo = 0, ghost_hi = 0;   // global row range owned / ghost ranges around it
    bool gathered = false;           // first level that lives on rank 0 only; every rank owns a slice of its RHS
    int64_t my_off = 0, my_cnt = 0;  // this rank's slice of the gathered level
    std::vector<int64_t> gather_off; // world + 1 offsets of all slices
    std::vector<int> peers, send_cnt, recv_cnt;
    int32_t* send_idx = nullptr;     // device: owned local indices to pack, peer after peer
    double* send_buf = nullptr;
    int64_t send_total = 0;
    // peer-memory halo exchange (CUDA IPC over NVLink): flags + two staging copies of the ghost section live in one
    // exported allocation; the neighbours write into it directly
    void* p2p_arena = nullptr;       // [flags: 256 x u64][fused flags: 256 x u64][stage 0: n_ghost][stage 1: n_ghost][v][vtmp]
    bool vec_in_arena = false;       // v and vtmp point into the arena (the neighbours store their boundary rows into them)
    int send_a[2] = {-1, -1};        // first row of the send list of neighbour p when the list is one run of consecutive rows
    bool fuse_ok = false;            // halo exchange fused into the kernels on this level (HaloFuse)
    unsigned long long* fuse_counters = nullptr;   // device: [p] epochs, [8 + p] CTA arrival counters
    unsigned long long* p2p_counters = nullptr;   // device: [0..15] send epochs, [16..31] recv epochs, [32] block counter x2
    std::vector<void*> p2p_opened;   // peers' arenas mapped into this process
    bool p2p_ready = false;
    int p2p_imported = 0;
    P2PPlan p2p;

    DevCsr A, RJ, P, R, G;           // G: Gauss-Seidel off-diagonal operator, rows in execution order
    double* dinv = nullptr;
    int32_t* inj = nullptr;
    std::vector<int64_t> perm_host;  // caller numbering (mgb_set_numbering): new index of every dof; kept for late operators (mass matrix)
    int32_t* perm = nullptr;         // ... on the device: vectors are permuted on the way in and out
    bool inj_mono = false;           // the injection list is strictly ascending
    int32_t* cmap = nullptr;         // fine dof -> coarse dof or -1 (fused residual + injection)
    int4* inj_desc = nullptr;        // the stream tiles of A that contain at least one injected row
    int inj_ntiles = 0;
    int inj_n_int = 0, inj_n_bnd = 0;   // sharded: the list is stored [interior tiles | boundary tiles]
    double inj_fraction = 1.0;       // share of A's entries in those tiles
    double *v = nullptr, *vtmp = nullptr, *f = nullptr, *r = nullptr, *g = nullptr;
    double* b = nullptr;             // re-discretised right-hand side b_dict[l] (FMG, multigrid.py:279)
    DevCsr M;                        // optional mass matrix for the L2(Omega) norm of the FMG stopping rule
    double* uex = nullptr;           // optional exact solution (nodal values): per-cycle error norms of the FMG driver
    // Gauss-Seidel artefacts (host copies are what mgb_get_artifact returns)
    std::vector<int32_t> lev_of_row, lev_order, lev_off, col_of_row, col_order, col_off;
    int32_t* gs_order = nullptr;     // device: execution order actually used by G
    int32_t* gs_off = nullptr;       // device: level offsets (GS_LEVEL)
    double* gs_diag = nullptr;       // device: a_ii in execution order
    int gs_groups = 0;               // number of levels / colours
    int gs_max_width = 0;
    int32_t* gs_ecols = nullptr;     // GS_LEVEL, rows of <= 8 entries: the operator again in ELL form (W x n, level-major)
    double* gs_evals = nullptr;
    int gs_W = 0;
    unsigned long long* s2 = nullptr;   // k_hotrow2 (two Jacobi sweeps per launch): first-sweep tiles finished per tile group
    int s2_groups = 0;
    int gs_setup_passes[2] = {0, 0}; // device set-up: relaxation passes the level sets / the colouring took
};

struct ProfEvent { int kind, level; double bytes, moved; cudaEvent_t e0, e1; };

}  // namespace

struct mgb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    std::map<int, Level> levels;
    bool finalized = false;
    double omega = 2.0 / 3.0;
    int mu1 = 2, mu2 = 2, smoother = MGB_SM_JACOBI_RJ;
    // options
    int rj_reversed = 1, use_graph = 1, opt_family = 0, opt_lpr = 0, opt_iter = 0, coarse_refine = 0, fuse_restrict = 1;
    int pdl = -1;                  // programmatic dependent launch: a kernel's barrier set-up, dictionary load and first matrix
                                   // tiles overlap the tail of its predecessor.  -1 (default): row-stream kernels of coded operators
                                   // only (measured faster); 1: the CSR stream kernels too (measured slower); 0: never
    int stream_auto = 0;           // pick the stream configuration per operator from its average row length (measured: no gain)
    int gs_cluster = 2;            // level-scheduled Gauss-Seidel: 0 grid barrier, 1 one cluster, 2 one cluster + ELL prefetch pipeline
    int stream_cfg = 3;            // 0: register-staged tile kernel; 1..6: TMA stream kernel configuration (stream_choice)
    int compress = 3;              // lossless coding of repetitive operators (mgb_code.cuh): 0 off, 1 per-entry codes, 2 + row patterns,
                                   // 3 + anchored row patterns for rectangular operators
    int code_cfg = 1;              // row-stream kernel configuration for coded operators (code_choice)
    int stage_x = 3;               // row-pattern-coded operators: 3 speculative loads at the hot pattern's offsets (k_hotrow, the default);
                                   // 1 x staged in shared memory by bulk copies (k_rowwin); 0 bulk-copied codes / operands + x gathered
                                   // through L1 (k_rowstream)
    int tail_rows = 0;             // > 0: levels of at most this many rows (and everything below them) run in one cooperative launch
                                   // (k_tail).  OFF by default: measured slower (profiles/r2_variants_tail_*.jsonl) -- inside a replayed
                                   // graph a small kernel costs ~2.2 us, a grid-wide barrier phase ~6 us (cfg2: 0.229 -> 0.307 ms)
    int tail_cluster = 0;          // k_tail as ONE 8-CTA cluster with the hardware cluster barrier between phases instead of the cooperative grid
                                   // (bit-identical, measured slower as well: cfg2 0.221 -> 0.278 ms)
    int reuse_g = 1;               // cycles after the first of one call reuse the top level's w*(dinv*f) instead of forming it again
    bool numbered = false;         // some level carries a caller numbering
    double* perm_tmp = nullptr;    // device scratch of the permuting copies
    int64_t perm_tmp_cap = 0;
    int device_setup = 0;          // 1: transposed restriction, level sets, colourings and Gauss-Seidel operators are built on the device for
                                   //    host-assembled levels too (generated levels always are); bit-identical to the host build
    int fuse_halo = 1;             // row-sharded levels: halo exchange fused into the kernels that write / read the iterate (HaloFuse)
    HaloFuse hf_cur{};             // the fused-exchange plan of the launch being enqueued (all zero: none)
    bool in_cycle = false;         // inside enqueue_cycle: ghost sections of fused levels are kept valid by the kernels themselves
    int hot_inj = 1;               // fused residual + injection: thread per coarse row on a pattern-coded level matrix (k_hotinj)
    int anch_cfg = 1;              // anchored-pattern kernel: 1 128 threads x 4 rows, 2 128 x 2, 3 64 x 4, 4 256 x 1 (rows of > 4 entries)
    int anch_tiles = 8;            // > 1: consecutive tiles per CTA of the anchored-pattern kernel on one GPU, the next tile's anchors / codes in flight
                                   //   (k_anchloop; 513^3 prolongation 0.91 -> 0.83 ms; 1: one tile per CTA, k_anchrow; negative: also on small grids)
    int hot_cfg = 1;               // hot-row kernel configuration (hot_choice)
    int fuse_sweeps = 0;           // Jacobi: 1 = pairs of sweeps in one launch on unsharded hot-row levels of at least s2_min_rows rows (k_hotrow2).
                                   //   OFF by default: bit-identical, DRAM traffic of a pair 6.7 -> 4.6 GB at 513^3, but no faster -- the sweep is
                                   //   bound by rows in flight x latency, not by bytes (profiles/r2_hotrow2_*.jsonl)
    int s2_slack = 3072;           //   tiles the second sweep trails the first by, beyond the reach of a row (+ one counter group)
    int s2_min_rows = 1 << 20;
    int s2_tiles = 4;              //   consecutive tiles per CTA (one release / one poll per CTA)
    int hot_pf = 262144;           // hot-row kernel: L2 prefetch distance in rows (0: none)
    int win_cfg = 1;               // row-window kernel configuration (win_choice)
    int win_prefetch = 0;          // row-window kernel: tiles (per CTA) whose DRAM streams are prefetched into L2 ahead of the copies
    bool allow_stream = true;      // false while borrowed user pointers are in play (no padding / alignment guarantee)
    int coarsest = 0, finest = 0;
    double* coarse_inv = nullptr;
    std::vector<double> coarse_inv_host;
    double *d_partial = nullptr, *d_hist = nullptr;
    int hist_cap = 0;
    int norm_blocks = 0;
    std::map<int, cudaGraphExec_t> graphs;
    std::map<int, int64_t> graph_kernels;
    std::vector<double> fmg_err;   // error norms of the last mgb_fmg run (one per finest-level cycle)
    bool prof = false;
    std::vector<ProfEvent> prof_events;
    std::map<std::pair<int, int>, mgb_profile_record> prof_records;
    int64_t launches = 0;
    // ---- multi-GPU: one process per GPU, NCCL communicator created from a broadcast unique id
    bool dist = false;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t comm_stream = nullptr;      // halo exchanges run here while interior rows run on `stream`
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int p2p_enable = 1;                      // option "p2p_enable": 0 forces ncclSend/ncclRecv even where peers are mapped
    int overlap = 0;                         // option "overlap_halo"
    int overlap_waves = 2;                   // option "overlap_waves": waves of retiring CTAs in an overlapped interior launch
    int gather_level = INT_MIN;      // level that is gathered to rank 0 (INT_MIN: none)
    int sm_count = 148;
    int gs_coop_blocks_per_sm = 0;
};

