// freemux.inl — freemuxlet kernels (stage 1, seeding, E-step, classify, M-step)
static void fmx_state_free(pscl_ctx* ctx) { (void)ctx; }
#define FMX_TODO(ctx) (ctx ? pscl_fail(ctx, PSCL_ESTATE, "freemuxlet path not built yet") : PSCL_EINVAL)
extern "C" int pscl_fmx_run(pscl_ctx* ctx, const pscl_pileup*, const pscl_fmx_opts*, const int32_t*, pscl_fmx_cell*, double*, int32_t*, pscl_fmx_result*) { return FMX_TODO(ctx); }
extern "C" int pscl_fmx_init(pscl_ctx* ctx, const pscl_plp*, const pscl_fmx_opts*) { return FMX_TODO(ctx); }
extern "C" int pscl_fmx_stage1(pscl_ctx* ctx, double*) { return FMX_TODO(ctx); }
extern "C" int pscl_fmx_seed(pscl_ctx* ctx, const double*, int32_t*) { return FMX_TODO(ctx); }
extern "C" int pscl_fmx_mstep(pscl_ctx* ctx, const int32_t*) { return FMX_TODO(ctx); }
extern "C" int pscl_fmx_estep(pscl_ctx* ctx, int32_t, double*) { return FMX_TODO(ctx); }
extern "C" int pscl_fmx_classify(pscl_ctx* ctx, const double*, int32_t*, pscl_fmx_result*) { return FMX_TODO(ctx); }
extern "C" int pscl_fmx_fetch(pscl_ctx* ctx, pscl_fmx_cell*, double*, int32_t*) { return FMX_TODO(ctx); }
extern "C" int pscl_fmx_last_kernel_ms(pscl_ctx* ctx, float*, float*, float*) { return FMX_TODO(ctx); }
