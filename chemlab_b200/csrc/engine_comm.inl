// engine_comm.inl -- multi-GPU slab decomposition over NCCL (included by engine.cu)
struct clb_engine::CommDev { int dummy; };

extern "C" int clb_nccl_unique_id(void* id128_out) {
    (void)id128_out;
    g_create_error = "multi-GPU support is not built yet";
    return CLB_ERR_UNSUPPORTED;
}
extern "C" int clb_comm_init(clb_engine* e, int rank, int nranks, const void* nccl_id128) {
    (void)rank; (void)nccl_id128;
    if (!e) return CLB_ERR_ARG;
    if (nranks == 1) return CLB_OK;
    return e->fail(CLB_ERR_UNSUPPORTED, "multi-GPU support is not built yet");
}
int clb_engine::comm_migrate_and_ghosts() { return CLB_OK; }
int clb_engine::comm_after_sort() { return CLB_OK; }
int clb_engine::comm_halo_positions() { return CLB_OK; }
int clb_engine::comm_allreduce_sum(double*, int) { return CLB_OK; }
int clb_engine::comm_gather_candidates(long long*) { return CLB_OK; }
void clb_engine::comm_destroy() {}
