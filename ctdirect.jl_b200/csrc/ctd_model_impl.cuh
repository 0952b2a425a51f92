// Included by every generated translation unit (csrc/generated/<element>.cu): instantiates the
// evaluation kernels for one element struct and registers them with the model registry.
#pragma once
#include "ctd_kernels.cuh"
#include "ctd_model.hpp"

namespace ctd {

// the output combinations that are compiled (one kernel each):
//   cons | jac | hess | obj | cons+obj | grad+obj | cons+jac+hess | everything
#define CTD_FOR_EACH_FLAGSET(X) X(1) X(2) X(4) X(16) X(17) X(24) X(7) X(31) X(49) X(63)
// 49 = cons+obj+violation, 63 = everything+violation: the per-instance violation norm reduced in the same
// launch; only compiled for light elements (the batched MPC use case), others use ctd_violation_kernel
constexpr bool viol_set(int f) { return (f & 32) != 0; }
// the "parts" variant (wide elements on small grids) is only compiled for the derivative-heavy sets
#define CTD_FOR_EACH_PARTS_FLAGSET(X) X(2) X(4) X(7) X(31)
constexpr bool parts_compiled(int f) { return f == 2 || f == 4 || f == 7 || f == 31; }

// launch with the programmatic-stream-serialization attribute (see pdl_entry() in ctd_kernels.cuh)
template <class K>
inline void launch_pdl(K kernel, dim3 grid, int threads, size_t smem, cudaStream_t st, const EvalArgs& a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = CTD_PDL;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, a);
}

template <class E>
struct ElementLauncher {
    template <int F>
    static LaunchCfg cfg_of() {
        using C = Cfg<E, F>;
        LaunchCfg c;
        c.threads = C::T;
        c.own = C::OWN;
        c.wpc = C::WPC;
        c.warp_smem = (size_t)C::WARP_DOUBLES * sizeof(double);
        c.nred = C::NREDX;
        c.tile_smem = C::TILE_SMEM;
        c.map_step = C::M_STEP + 1;
        c.map_fp = C::M_FP;
        c.map_bc = C::M_BC;
        if constexpr (!C::LIGHT && parts_compiled(F)) {
            using PC = PCfg<E, F>;
            c.p_threads = PC::T;
            c.p_own = PC::OWN;
            c.p_smem = PC::TILE_SMEM;
        }
        return c;
    }
    static LaunchCfg cfg(int flags) {
        switch (flags) {
#define CTD_X(F) case F: if constexpr (viol_set(F) && !Cfg<E, F>::LIGHT) return LaunchCfg(); else return cfg_of<F>();
            CTD_FOR_EACH_FLAGSET(CTD_X)
#undef CTD_X
        }
        return LaunchCfg();
    }
    static int prepare(int flags, size_t smem, int parts) {
        if (parts) {
            switch (flags) {
#define CTD_X(F) \
    case F: if constexpr (!Cfg<E, F>::LIGHT) return (int)cudaFuncSetAttribute(ctd_eval_kernel_parts<E, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); else return (int)cudaErrorInvalidValue;
                CTD_FOR_EACH_PARTS_FLAGSET(CTD_X)
#undef CTD_X
            }
            return (int)cudaErrorInvalidValue;
        }
        switch (flags) {
#define CTD_X(F) \
    case F: if constexpr (Cfg<E, F>::LIGHT) return (int)cudaFuncSetAttribute(ctd_eval_kernel_light<E, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            else if constexpr (viol_set(F)) return (int)cudaErrorInvalidValue; \
            else return (int)cudaFuncSetAttribute(ctd_eval_kernel<E, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            CTD_FOR_EACH_FLAGSET(CTD_X)
#undef CTD_X
        }
        return (int)cudaErrorInvalidValue;
    }
    static int launch(const EvalArgs& a, int flags, int batch, size_t smem, cudaStream_t st, int parts, int threads) {
        dim3 grid((unsigned)a.ctas, (unsigned)batch, 1);
        if (parts) {
            switch (flags) {
#define CTD_X(F) \
    case F: if constexpr (!Cfg<E, F>::LIGHT) launch_pdl(ctd_eval_kernel_parts<E, F>, grid, PCfg<E, F>::T, smem, st, a); break;
                CTD_FOR_EACH_PARTS_FLAGSET(CTD_X)
#undef CTD_X
                default: return (int)cudaErrorInvalidValue;
            }
            return (int)cudaGetLastError();
        }
        switch (flags) {
#define CTD_X(F) \
    case F: if constexpr (Cfg<E, F>::LIGHT) launch_pdl(ctd_eval_kernel_light<E, F>, grid, threads, smem, st, a); \
            else if constexpr (viol_set(F)) return (int)cudaErrorInvalidValue; \
            else launch_pdl(ctd_eval_kernel<E, F>, grid, threads, smem, st, a); break;
            CTD_FOR_EACH_FLAGSET(CTD_X)
#undef CTD_X
            default: return (int)cudaErrorInvalidValue;
        }
        return (int)cudaGetLastError();
    }
};

template <class E>
inline int launch_unpack(const UnpackArgs& a, int batch, cudaStream_t st) {
    dim3 grid((unsigned)((a.N + 1 + 127) / 128), (unsigned)batch, 1);
    ctd_unpack_kernel<E><<<grid, 128, 0, st>>>(a);
    return (int)cudaGetLastError();
}

template <class E>
inline int launch_pack(const PackArgs& a, int batch, cudaStream_t st) {
    dim3 grid((unsigned)((a.N + 1 + 127) / 128), (unsigned)batch, 1);
    ctd_pack_kernel<E><<<grid, 128, 0, st>>>(a);
    return (int)cudaGetLastError();
}

template <class E>
struct ElementRegistrar {
    ElementRegistrar() {
        Registration r;
        r.info = make_elem_info<E>();
        r.launcher.launch = &ElementLauncher<E>::launch;
        r.launcher.cfg = &ElementLauncher<E>::cfg;
        r.launcher.prepare = &ElementLauncher<E>::prepare;
        r.launcher.unpack = &launch_unpack<E>;
        r.launcher.pack = &launch_pack<E>;
        register_element(r);
    }
};

}  // namespace ctd

#define CTD_REGISTER_ELEMENT(E) \
    namespace { ctd::ElementRegistrar<E> ctd_registrar_instance; }
