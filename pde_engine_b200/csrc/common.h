// common.h -- shared host-side definitions of libpde_b200 (error handling, handles).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/pde_b200.h"

namespace pde {

void set_error(const char* fmt, ...);
int cuda_fail(int err, const char* what);   // records message, returns PDE_E_CUDA
void count_launch(int n = 1);
bool have_device();
// Stream-ordered scratch memory from a pool the library owns (one per device, never trimmed): per-call work space
// without file-scope buffers -- safe with several host threads and streams -- and without paying cudaMalloc /
// cudaFree around every call (8-70 ms for the 400 MB dedup table).
int scratch_alloc(void** ptr, size_t bytes, void* stream);
void scratch_free(void* ptr, void* stream);
// order.cu: evaluation order of a batch for the stage-2 kernel -- candidate indices sorted by the leading program
// bytes, so that the warp groups resident on an SM run near-identical micro-op streams (instruction-cache locality).
// *order comes from scratch_alloc (release with scratch_free on the same stream); null when the batch is too small.
int candidate_order(const uint8_t* code, const unsigned* row_off, const uint8_t* len, long long n, int L, int** order, void* stream);
void exprset_mark_use(const pde_exprset* e, void* stream);     // pde_b200.cu: the handle's mirrors are in use on `stream` up to here

#define PDE_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t _e = (call);                                           \
        if (_e != cudaSuccess) return ::pde::cuda_fail((int)_e, #call);    \
    } while (0)

}  // namespace pde

struct pde_exprset;
namespace pde {
void exprset_ensure_rank(pde_exprset* e);      // compiler.cpp
int exprset_ensure_device(pde_exprset* e);     // compiler.cpp
}

// ---- opaque handles ------------------------------------------------------
struct pde_session {
    std::string var[2];
    std::vector<std::string> named;       // names of symbolic constants
    std::vector<double> named_vals;
    std::vector<std::string> const_keys;  // slot -> key
    std::vector<double> const_vals;
    std::vector<std::string> pow_keys;
    std::vector<double> pow_vals;
    // numeric mirrors of the keys (numerator, denominator; denominator 0 = named constant index):
    // rebuilt by the compiler whenever their size differs from the key tables
    std::vector<long long> const_num, const_den, pow_num, pow_den;
};

struct pde_exprset {
    int n = 0;
    std::vector<uint8_t> flags, attrs;
    std::vector<uint32_t> rank;
    std::vector<uint32_t> term_begin;   // [n+1]
    std::vector<int8_t> term_sign;      // [nt]
    std::vector<uint32_t> term_off;     // [nt+1]
    std::vector<uint8_t> pool;
    std::vector<char> str_blob;         // the source strings (NUL separated): lazy lexicographic rank
    std::vector<uint32_t> str_off;      // [n+1]
    bool rank_ready = false;
    // device mirrors (from the library's stream-ordered pool, on first use by the enumerator; cudaMalloc / cudaFree
    // per handle cost 1-20 ms and, now and then, several hundred -- more than a whole depth-4 validation)
    int device = -1;
    void* used_event = nullptr;         // cudaEvent_t: recorded after the last kernel that reads the mirrors; the free waits for it
    uint8_t* d_flags = nullptr;
    uint8_t* d_attrs = nullptr;
    uint32_t* d_rank = nullptr;
    uint32_t* d_term_begin = nullptr;
    int8_t* d_term_sign = nullptr;
    uint32_t* d_term_off = nullptr;
    uint8_t* d_pool = nullptr;
    // result of the enumerator's count pass (enumerate.cu), cached on the handle: a windowed pde_enumerate (one window
    // per rank / per chunk) does not repeat it.  Valid for (count_depth, count_prune, count_db) on `device`.
    int count_depth = 0, count_prune = -1;
    int32_t count_db[16] = {0};
    long long count_blocks = 0, count_total = 0;
    unsigned* d_count_sums = nullptr;             // candidates per block
    unsigned* d_count_in_tile = nullptr;          // exclusive prefix inside a 1024-block tile
    unsigned long long* d_count_tile = nullptr;   // exclusive prefix of the tile totals
    std::vector<unsigned long long> count_tile_host;   // the same on the host: a window launches only its tiles
    // CSR form (pde_enumerate_csr): bytes per block for row length L = count_bytes_L, scanned the same way, and the
    // per-block exclusive prefixes of candidates and bytes on the host (a window's pool range needs no device read)
    int count_bytes_L = 0;
    unsigned* d_bytes_sums = nullptr;
    unsigned* d_bytes_in_tile = nullptr;
    unsigned long long* d_bytes_tile = nullptr;
    std::vector<unsigned long long> block_cand_host, block_bytes_host;     // [nblocks + 1]
    std::vector<uint32_t> desc;         // [n][2] splice descriptors (enumerate.cu)
    std::vector<uint8_t> wpool;         // whole programs, padded by 8 bytes
    uint32_t* d_desc = nullptr;
    uint8_t* d_wpool = nullptr;
};

struct pde_program {
    int problem = 0;
    int order = 0;
    int n_coef = 0;
    int cols = 0;
    double consts[4] = {0, 0, 0, 0};
    // PDE_PROBLEM_PROGRAM: the run-time residual program (pde_compile_residual_program)
    std::vector<uint32_t> words;
    std::vector<double> prog_consts;
    int n_file = 0;
};

// program.cu: stage 2 with a run-time residual program (its own instantiations of the kernel and its own tables)
struct pde_validate_out;
namespace pde {
struct ValidateParams;
int program_validate(const pde_session* s, const pde_program* p, const ValidateParams& vp, const pde_validate_out* out,
                     int confirm_points, double tau, double t0, void* stream);
int program_eval_points(const pde_session* s, const pde_program* p, const ValidateParams& vp, double tau, double t0, void* stream);
}
