// program.cu -- stage 2 with a RUN-TIME residual program (pde_compile_residual_program, include/pde_b200.h):
// the instantiations of the interpreter kernel whose residual operator is the program interpreter of validate.cuh
// (ResidualProg<2>, ResidualProg<4>), kept in their own translation unit so they compile in parallel with the
// built-in specialisations and own their copy of the __constant__ tables.
#include <cuda_runtime.h>
#include <string.h>
#include <vector>

#include "common.h"
#include "validate.cuh"
#include "launch.cuh"

namespace pde {

// One CTA per SM of 12 warps: the residual's file lives in the lane's spill column, and 12 warps leave room for three
// spill slots (51 entries at order 4) next to the staged programs at L = 128; 4-warp CTAs for small grids / DUMP.
template <int PROBLEM, bool DUMP, bool MAJ>
static int launch_program(const ValidateParams& vp, cudaStream_t st) {
    if constexpr (DUMP) {
        return launch_validate_cfg<PROBLEM, DUMP, 4, 1, 4, MAJ>(vp, st, nullptr);
    } else {
        bool fits = false;
        if (vp.P_eval >= 128) {
            int rc = launch_validate_cfg<PROBLEM, DUMP, 12, 1, 1, MAJ>(vp, st, &fits);
            if (rc || fits) return rc;
        }
        return launch_validate_cfg<PROBLEM, DUMP, 4, 1, 4, MAJ>(vp, st, nullptr);
    }
}

template <int PROBLEM>
struct ProgramLauncher {
    template <bool MAJ> static int launch(const ValidateParams& vp, cudaStream_t st) { return launch_program<PROBLEM, false, MAJ>(vp, st); }
};

// the program's words / constants / dimensions -> this unit's constant bank; the file must fit the spill column
static int upload_program(const pde_program* p, ValidateParams& vp, cudaStream_t st) {
    std::vector<uint32_t> wbuf(kResMaxWords, 0u);            // PDE_R_END padding (per call: no shared host state)
    uint32_t* words = wbuf.data();
    memcpy(words, p->words.data(), sizeof(uint32_t) * p->words.size());
    double consts[kResMaxConsts] = {0};
    for (size_t i = 0; i < p->prog_consts.size(); ++i) consts[i] = p->prog_consts[i];
    const int dims[2] = {p->cols, (int)p->prog_consts.size()};
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_res_words, words, sizeof(uint32_t) * (p->words.size() + 1 < (size_t)kResMaxWords ? p->words.size() + 1 : (size_t)kResMaxWords), 0, cudaMemcpyHostToDevice, st));
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_res_consts, consts, sizeof(consts), 0, cudaMemcpyHostToDevice, st));
    PDE_CUDA(cudaMemcpyToSymbolAsync(c_res_dims, dims, sizeof(dims), 0, cudaMemcpyHostToDevice, st));
    // (pageable host sources: cudaMemcpy*Async returns once they are staged, like the table uploads of launch.cuh)
    // the spill column doubles as the residual's file: ns slots of (n_coef + 2) entries per lane
    const int per_slot = p->n_coef + 2;
    const int need = (p->n_file + per_slot - 1) / per_slot;
    if (need > vp.ns) vp.ns = need;
    if (vp.ns > 8) { set_error("residual program file of %d entries does not fit the spill column", p->n_file); return PDE_E_OVERFLOW; }
    return PDE_OK;
}

int program_validate(const pde_session* s, const pde_program* p, const ValidateParams& vp_in, const pde_validate_out* out,
                     int confirm_points, double tau, double t0, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ValidateParams vp = vp_in;
    int rc = upload_tables(s, tau, t0, st);
    if (rc) return rc;
    rc = upload_program(p, vp, st);
    if (rc) return rc;
    return p->order == 4 ? run_two_pass<ProgramLauncher<kProblemProgram4>>(vp, out, confirm_points, st)
                         : run_two_pass<ProgramLauncher<kProblemProgram2>>(vp, out, confirm_points, st);
}

int program_eval_points(const pde_session* s, const pde_program* p, const ValidateParams& vp_in, double tau, double t0, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ValidateParams vp = vp_in;
    int rc = upload_tables(s, tau, t0, st);
    if (rc) return rc;
    rc = upload_program(p, vp, st);
    if (rc) return rc;
    return p->order == 4 ? launch_program<kProblemProgram4, true, true>(vp, st) : launch_program<kProblemProgram2, true, true>(vp, st);
}

}  // namespace pde
