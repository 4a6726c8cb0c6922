// host_pipeline.cu -- ct_host_run: the op on HOST buffers.  Rows are streamed through the device
// in chunks; each chunk does H2D -> kernel -> D2H on its slot's stream, and NSLOT slots rotate so
// that the upload of chunk k+1, the kernel of chunk k and the download of chunk k-1 overlap on the
// two copy engines and the SMs.  Scales / zero points are uploaded once.  Device scratch is cached
// per device (grow-only) and serialised by a mutex.
#include <mutex>

#include "engine.h"

namespace ctb {

constexpr int NSLOT = 3;

struct Scratch {
    std::mutex mu;
    void* in[NSLOT] = {nullptr, nullptr, nullptr};
    void* out[NSLOT] = {nullptr, nullptr, nullptr};
    size_t in_cap = 0, out_cap = 0;
    void* aux = nullptr;   // scale | zp
    size_t aux_cap = 0;
    cudaStream_t st[NSLOT] = {nullptr, nullptr, nullptr};
    bool init = false;
};
static Scratch g_scratch[16];

static int ensure(Scratch& s, size_t in_bytes, size_t out_bytes, size_t aux_bytes) {
    if (!s.init) {
        for (int i = 0; i < NSLOT; ++i) CT_CUDA_TRY(cudaStreamCreateWithFlags(&s.st[i], cudaStreamNonBlocking));
        s.init = true;
    }
    if (in_bytes > s.in_cap) {
        for (int i = 0; i < NSLOT; ++i) {
            if (s.in[i]) cudaFree(s.in[i]);
            CT_CUDA_TRY(cudaMalloc(&s.in[i], in_bytes));
        }
        s.in_cap = in_bytes;
    }
    if (out_bytes > s.out_cap) {
        for (int i = 0; i < NSLOT; ++i) {
            if (s.out[i]) cudaFree(s.out[i]);
            CT_CUDA_TRY(cudaMalloc(&s.out[i], out_bytes));
        }
        s.out_cap = out_bytes;
    }
    if (aux_bytes > s.aux_cap) {
        if (s.aux) cudaFree(s.aux);
        CT_CUDA_TRY(cudaMalloc(&s.aux, aux_bytes));
        s.aux_cap = aux_bytes;
    }
    return CT_OK;
}

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace ctb

extern "C" int ct_host_run(int op, const ct_quant_desc* d, const void* in, const void* scale, const void* zp, void* out, int device) {
    using namespace ctb;
    if (!d) { set_error("null descriptor"); return CT_E_ARG; }
    int rc = check_device(device);
    if (rc) return rc;
    if (device >= 16) { set_error("device index too large"); return CT_E_ARG; }
    if (d->rows * d->cols == 0) return CT_OK;
    if (!in || !scale || !out) { set_error("null tensor pointer"); return CT_E_ARG; }
    DeviceGuard guard(device);

    // bytes per row of the streamed input / output
    const int64_t nw = (d->cols * d->num_bits + 31) / 32;
    size_t in_row, out_row;
    switch (op) {
    case CT_OP_QUANTIZE_PACK: in_row = (size_t)d->cols * dt_size(d->x_dtype); out_row = (size_t)nw * 4; break;
    case CT_OP_UNPACK_DEQUANTIZE: in_row = (size_t)nw * 4; out_row = (size_t)d->cols * dt_size(d->out_dtype); break;
    case CT_OP_QUANTIZE: in_row = (size_t)d->cols * dt_size(d->x_dtype); out_row = (size_t)d->cols * dt_size(d->q_dtype); break;
    case CT_OP_DEQUANTIZE: in_row = (size_t)d->cols * dt_size(d->q_dtype); out_row = (size_t)d->cols * dt_size(d->out_dtype); break;
    case CT_OP_FAKE_QUANTIZE: in_row = (size_t)d->cols * dt_size(d->x_dtype); out_row = (size_t)d->cols * dt_size(d->out_dtype); break;
    default: set_error("unknown op %d", op); return CT_E_ARG;
    }
    if (in_row == 0 || out_row == 0) { set_error("bad dtype in descriptor"); return CT_E_DTYPE; }

    // scale / zero-point extents
    const bool row_scaled = (d->rdiv != CT_DIV_INF);
    const int64_t row_blocks = row_scaled ? (d->rows + d->rdiv - 1) / d->rdiv : 1;
    const int64_t cols_per_row = (d->cdiv == CT_DIV_INF) ? 1 : (d->cols + d->cdiv - 1) / d->cdiv;
    const int64_t n_scale = (row_scaled && d->s_row_stride > 0) ? row_blocks * d->s_row_stride : cols_per_row;
    const size_t s_bytes = round_up((size_t)n_scale * dt_size(d->scale_dtype), 256);
    const size_t z_bytes = zp ? round_up((size_t)n_scale * dt_size(d->zp_dtype), 256) : 0;

    // chunk: ~32 MiB of input, a multiple of the row block
    int64_t rows_per_chunk = (int64_t)((32u << 20) / in_row);
    if (rows_per_chunk < 1) rows_per_chunk = 1;
    if (row_scaled && d->rdiv > 1) rows_per_chunk = (rows_per_chunk + d->rdiv - 1) / d->rdiv * d->rdiv;
    if (rows_per_chunk > d->rows) rows_per_chunk = d->rows;

    Scratch& S = g_scratch[device];
    std::lock_guard<std::mutex> lk(S.mu);
    rc = ensure(S, (size_t)rows_per_chunk * in_row, (size_t)rows_per_chunk * out_row, s_bytes + z_bytes);
    if (rc) return rc;

    uint8_t* d_scale = reinterpret_cast<uint8_t*>(S.aux);
    uint8_t* d_zp = zp ? d_scale + s_bytes : nullptr;
    CT_CUDA_TRY(cudaMemcpyAsync(d_scale, scale, (size_t)n_scale * dt_size(d->scale_dtype), cudaMemcpyHostToDevice, S.st[0]));
    if (zp) CT_CUDA_TRY(cudaMemcpyAsync(d_zp, zp, (size_t)n_scale * dt_size(d->zp_dtype), cudaMemcpyHostToDevice, S.st[0]));
    cudaEvent_t aux_ready;
    CT_CUDA_TRY(cudaEventCreateWithFlags(&aux_ready, cudaEventDisableTiming));
    CT_CUDA_TRY(cudaEventRecord(aux_ready, S.st[0]));
    for (int i = 1; i < NSLOT; ++i) CT_CUDA_TRY(cudaStreamWaitEvent(S.st[i], aux_ready, 0));

    int slot = 0;
    for (int64_t r0 = 0; r0 < d->rows; r0 += rows_per_chunk, slot = (slot + 1) % NSLOT) {
        const int64_t nr = (d->rows - r0 < rows_per_chunk) ? d->rows - r0 : rows_per_chunk;
        cudaStream_t st = S.st[slot];
        CT_CUDA_TRY(cudaMemcpyAsync(S.in[slot], reinterpret_cast<const uint8_t*>(in) + (size_t)r0 * in_row, (size_t)nr * in_row, cudaMemcpyHostToDevice, st));
        ct_quant_desc sub = *d;
        sub.rows = nr;
        const int64_t sblock = row_scaled ? (r0 / d->rdiv) * d->s_row_stride : 0;
        const void* sc = d_scale + (size_t)sblock * dt_size(d->scale_dtype);
        const void* zz = d_zp ? d_zp + (size_t)sblock * dt_size(d->zp_dtype) : nullptr;
        const void* ins[1] = {S.in[slot]};
        const void* scs[1] = {sc};
        const void* zps[1] = {zz};
        void* outs[1] = {S.out[slot]};
        rc = run_batched(op, 1, &sub, ins, scs, zps, nullptr, outs, device, st);
        if (rc) { cudaEventDestroy(aux_ready); return rc; }
        CT_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out) + (size_t)r0 * out_row, S.out[slot], (size_t)nr * out_row, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < NSLOT; ++i) CT_CUDA_TRY(cudaStreamSynchronize(S.st[i]));
    cudaEventDestroy(aux_ready);
    return CT_OK;
}
