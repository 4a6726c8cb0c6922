// host_many.cu -- ct_host_run_many: one op over MANY host-resident tensors (a CPU-resident model),
// pipelined ACROSS tensors.  All row chunks of all tensors form one queue that is fed through NSLOT
// staging slots (own stream each: H2D chunk -> kernel -> D2H chunk), so the H2D copy engine never
// drains between tensors; every tensor's scales / zero points are uploaded once into a per-call aux
// area and published to the other slots with an event.  Blocking: returns when all outputs are in
// host memory.  (ct_host_run, one tensor per call, restarts the pipeline for every tensor.)
#include <cstdlib>
#include <mutex>
#include <vector>

#include "engine.h"

namespace ctb {

constexpr int MSLOT = 8;   // slots allocated; CT_B200_HOST_SLOTS (default 4) of them are used
static int host_slots() { const char* v = getenv("CT_B200_HOST_SLOTS"); int n = v ? atoi(v) : 4; return n < 2 ? 2 : (n > MSLOT ? MSLOT : n); }
static size_t host_chunk() { const char* v = getenv("CT_B200_HOST_CHUNK_MB"); int n = v ? atoi(v) : 32; return (size_t)(n < 1 ? 1 : (n > 512 ? 512 : n)) << 20; }

struct ManyScratch {
    std::mutex mu;
    void* in[MSLOT] = {};
    void* out[MSLOT] = {};
    size_t in_cap = 0, out_cap = 0;
    void* aux = nullptr;
    size_t aux_cap = 0;
    cudaStream_t st[MSLOT] = {};
    bool init = false;
};
static ManyScratch g_many[16];

static size_t up256(size_t v) { return (v + 255) / 256 * 256; }

static int row_bytes(int op, const ct_quant_desc& d, size_t& in_row, size_t& out_row) {
    const int64_t nw = (d.cols * d.num_bits + 31) / 32;
    switch (op) {
    case CT_OP_QUANTIZE_PACK: in_row = (size_t)d.cols * dt_size(d.x_dtype); out_row = (size_t)nw * 4; break;
    case CT_OP_UNPACK_DEQUANTIZE: in_row = (size_t)nw * 4; out_row = (size_t)d.cols * dt_size(d.out_dtype); break;
    case CT_OP_QUANTIZE: in_row = (size_t)d.cols * dt_size(d.x_dtype); out_row = (size_t)d.cols * dt_size(d.q_dtype); break;
    case CT_OP_DEQUANTIZE: in_row = (size_t)d.cols * dt_size(d.q_dtype); out_row = (size_t)d.cols * dt_size(d.out_dtype); break;
    case CT_OP_FAKE_QUANTIZE: in_row = (size_t)d.cols * dt_size(d.x_dtype); out_row = (size_t)d.cols * dt_size(d.out_dtype); break;
    case CT_OP_QUANTIZE_PACK_FP4: in_row = (size_t)d.cols * dt_size(d.x_dtype); out_row = (size_t)d.cols / 2; break;
    case CT_OP_UNPACK_DEQUANTIZE_FP4: in_row = (size_t)d.cols / 2; out_row = (size_t)d.cols * dt_size(d.out_dtype); break;
    default: set_error("unknown op %d", op); return CT_E_ARG;
    }
    if (in_row == 0 || out_row == 0) { set_error("bad dtype in descriptor"); return CT_E_DTYPE; }
    return CT_OK;
}

static int64_t scale_count(const ct_quant_desc& d) {
    const bool row_scaled = (d.rdiv != CT_DIV_INF);
    const int64_t row_blocks = row_scaled ? (d.rows + d.rdiv - 1) / d.rdiv : 1;
    const int64_t per_row = (d.cdiv == CT_DIV_INF) ? 1 : (d.cols + d.cdiv - 1) / d.cdiv;
    return (row_scaled && d.s_row_stride > 0) ? row_blocks * d.s_row_stride : per_row;
}

}  // namespace ctb

extern "C" int ct_host_run_many(int op, int n, const ct_quant_desc* descs, const void* const* in, const void* const* scale,
                                const void* const* zp, void* const* out, int device) {
    using namespace ctb;
    if (n < 0 || (n > 0 && (!descs || !in || !scale || !out))) { set_error("null table"); return CT_E_ARG; }
    int rc = check_device(device);
    if (rc) return rc;
    if (device >= 16) { set_error("device index too large"); return CT_E_ARG; }
    if (n == 0) return CT_OK;
    DeviceGuard guard(device);

    const size_t CHUNK = host_chunk();   // bytes of streamed input per chunk
    const int NS = host_slots();
    struct T { size_t in_row, out_row, s_off, z_off, s_bytes, z_bytes, g_off; int64_t rpc; };   // g_off: the tensor's global scale (HOST pointer in
                                                                                                // the descriptor of this entry point), uploaded to aux
    std::vector<T> ts((size_t)n);
    size_t aux_total = 0, max_in = 0, max_out = 0;
    for (int i = 0; i < n; ++i) {
        const ct_quant_desc& d = descs[i];
        T& t = ts[i];
        if (d.rows * d.cols == 0) { t = T{}; continue; }
        if (!in[i] || !scale[i] || !out[i]) { set_error("null tensor pointer (tensor %d)", i); return CT_E_ARG; }
        rc = row_bytes(op, d, t.in_row, t.out_row);
        if (rc) return rc;
        const int64_t ns = scale_count(d);
        t.s_bytes = (size_t)ns * dt_size(d.scale_dtype);
        t.z_bytes = (zp && zp[i]) ? (size_t)ns * dt_size(d.zp_dtype) : 0;
        t.s_off = aux_total; aux_total += up256(t.s_bytes);
        t.z_off = aux_total; aux_total += up256(t.z_bytes);
        t.g_off = aux_total; aux_total += d.global_scale ? 256 : 0;
        int64_t rpc = (int64_t)(CHUNK / t.in_row);
        if (rpc < 1) rpc = 1;
        if (d.rdiv != CT_DIV_INF && d.rdiv > 1) rpc = (rpc + d.rdiv - 1) / d.rdiv * d.rdiv;
        if (rpc > d.rows) rpc = d.rows;
        t.rpc = rpc;
        if ((size_t)rpc * t.in_row > max_in) max_in = (size_t)rpc * t.in_row;
        if ((size_t)rpc * t.out_row > max_out) max_out = (size_t)rpc * t.out_row;
    }

    ManyScratch& S = g_many[device];
    std::lock_guard<std::mutex> lk(S.mu);
    if (!S.init) {
        for (int i = 0; i < MSLOT; ++i) CT_CUDA_TRY(cudaStreamCreateWithFlags(&S.st[i], cudaStreamNonBlocking));
        S.init = true;
    }
    if (max_in > S.in_cap) {
        for (int i = 0; i < MSLOT; ++i) { if (S.in[i]) cudaFree(S.in[i]); CT_CUDA_TRY(cudaMalloc(&S.in[i], max_in)); }
        S.in_cap = max_in;
    }
    if (max_out > S.out_cap) {
        for (int i = 0; i < MSLOT; ++i) { if (S.out[i]) cudaFree(S.out[i]); CT_CUDA_TRY(cudaMalloc(&S.out[i], max_out)); }
        S.out_cap = max_out;
    }
    if (aux_total > S.aux_cap) {
        if (S.aux) cudaFree(S.aux);
        CT_CUDA_TRY(cudaMalloc(&S.aux, aux_total));
        S.aux_cap = aux_total;
    }
    uint8_t* aux = reinterpret_cast<uint8_t*>(S.aux);

    std::vector<cudaEvent_t> evs;
    int slot = 0;
    for (int i = 0; i < n; ++i) {
        const ct_quant_desc& d = descs[i];
        const T& t = ts[i];
        if (d.rows * d.cols == 0) continue;
        // qparams of this tensor: uploaded on the slot of its first chunk, published to the others
        cudaStream_t s0 = S.st[slot];
        CT_CUDA_TRY(cudaMemcpyAsync(aux + t.s_off, scale[i], t.s_bytes, cudaMemcpyHostToDevice, s0));
        if (t.z_bytes) CT_CUDA_TRY(cudaMemcpyAsync(aux + t.z_off, zp[i], t.z_bytes, cudaMemcpyHostToDevice, s0));
        if (d.global_scale) CT_CUDA_TRY(cudaMemcpyAsync(aux + t.g_off, d.global_scale, sizeof(float), cudaMemcpyHostToDevice, s0));
        cudaEvent_t ev;
        CT_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CT_CUDA_TRY(cudaEventRecord(ev, s0));
        evs.push_back(ev);
        const bool row_scaled = (d.rdiv != CT_DIV_INF);
        bool first = true;
        for (int64_t r0 = 0; r0 < d.rows; r0 += t.rpc, slot = (slot + 1) % NS) {
            const int64_t nr = (d.rows - r0 < t.rpc) ? d.rows - r0 : t.rpc;
            cudaStream_t st = S.st[slot];
            if (!first) CT_CUDA_TRY(cudaStreamWaitEvent(st, ev, 0));
            first = false;
            CT_CUDA_TRY(cudaMemcpyAsync(S.in[slot], reinterpret_cast<const uint8_t*>(in[i]) + (size_t)r0 * t.in_row, (size_t)nr * t.in_row, cudaMemcpyHostToDevice, st));
            ct_quant_desc sub = d;
            sub.rows = nr;
            if (d.global_scale) sub.global_scale = aux + t.g_off;
            const int64_t sblock = row_scaled ? (r0 / d.rdiv) * d.s_row_stride : 0;
            const void* sc = aux + t.s_off + (size_t)sblock * dt_size(d.scale_dtype);
            const void* zz = t.z_bytes ? aux + t.z_off + (size_t)sblock * dt_size(d.zp_dtype) : nullptr;
            const void* ins[1] = {S.in[slot]};
            const void* scs[1] = {sc};
            const void* zps[1] = {zz};
            void* outs[1] = {S.out[slot]};
            rc = run_batched(op, 1, &sub, ins, scs, zps, nullptr, outs, device, st);
            if (rc) { for (auto e : evs) cudaEventDestroy(e); return rc; }
            CT_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out[i]) + (size_t)r0 * t.out_row, S.out[slot], (size_t)nr * t.out_row, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int i = 0; i < MSLOT; ++i) CT_CUDA_TRY(cudaStreamSynchronize(S.st[i]));
    for (auto e : evs) cudaEventDestroy(e);
    return CT_OK;
}
