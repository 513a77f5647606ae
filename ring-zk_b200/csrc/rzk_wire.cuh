// rzk_wire.cuh -- the reference's wire format for batches of messages, packed and parsed on the device.
//
// The crate derives serde::Serialize / Deserialize for its message structs (commit.rs:134-141, 222-235;
// prove/open.rs:180-228, prove/linear.rs:256-315, prove/sum.rs:327-391) and its only serde test uses bincode 1.3.3
// (Cargo.toml dev-dependencies; mat.rs:424-438): little endian, fixed-width integers, a u64 length in front of every
// sequence, one tag byte in front of an Option, struct fields in declaration order with no framing.  A Mat is
// Vec<Vec<Polynomial>> (mat.rs:11-17), and a Polynomial serialises as the sequence of its stored coefficients:
// mat.rs:434 pins 8 (rows) + 8 (columns) + 8 (coefficients) + 3 * 4 = 36 bytes for the 1 x 1 matrix [1 + 2X + 3X^2] over i32.
//
// Everything a message contains is therefore a walk over a fixed list of tokens -- literal u64 lengths, Option tags and
// polynomials -- which the host builds once per message kind (rzk_wire_layout) and one warp per batch item replays:
//   pack:    sizes (trimmed polynomials make items differ in length) -> exclusive scan -> bytes
//   unpack:  bytes + item offsets -> coefficient arrays, every literal checked, malformed items flagged
// Two facts about the dependency's Polynomial / ZqI64 serde cannot be checked here (poly-ring-xnp1 is absent) and are
// parameters instead of assumptions: whether trailing zero coefficients are stored (trim = 1 reproduces the 36-byte
// case; trim = 0 always writes N coefficients) and the width of one coefficient (8 bytes for ZqI64's i64, 4 for i32).
#pragma once

namespace {

constexpr int kWireMaxStreams = 8;

struct WireLaunch {
    const rzk_wire_tok *toks;
    uint32_t ntoks;
    uint32_t n_items;
    const void *base[kWireMaxStreams];      // pack: inputs; unpack: outputs
    uint32_t polys[kWireMaxStreams];
    uint32_t dtype[kWireMaxStreams];        // DT_I32 / DT_I8
    uint32_t elem_bytes, trim;
    uint8_t *bytes;                         // packed messages
    uint64_t *sizes;                        // pack pass 1: bytes of every item
    const uint64_t *offsets;                // [n_items + 1]
    uint64_t in_bytes;                      // unpack: length of `bytes`
    uint32_t *flags;                        // unpack: FLAG_FAIL = malformed item
    int64_t q;
};

__device__ __forceinline__ int32_t wire_coeff(const WireLaunch &K, uint32_t s, uint64_t poly, uint32_t i)
{
    return K.dtype[s] == DT_I8 ? (int32_t)reinterpret_cast<const int8_t *>(K.base[s])[poly * kN + i]
                               : reinterpret_cast<const int32_t *>(K.base[s])[poly * kN + i];
}

// number of stored coefficients of a polynomial: N, or (trim) the index of the highest non-zero coefficient + 1
__device__ __forceinline__ uint32_t wire_poly_len(const WireLaunch &K, uint32_t s, uint64_t poly, uint32_t lane)
{
    if (!K.trim) return (uint32_t)kN;
    uint32_t hi = 0;
#pragma unroll 4
    for (int j = 0; j < kN / 32; ++j) {
        const uint32_t i = lane + 32u * (uint32_t)j;
        if (wire_coeff(K, s, poly, i) != 0) hi = i + 1u;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    return hi;
}

__device__ __forceinline__ void wire_put(uint8_t *p, uint64_t v, uint32_t nbytes)       // little endian, any alignment
{
    if (nbytes == 8 && (reinterpret_cast<uintptr_t>(p) & 7) == 0) { *reinterpret_cast<uint64_t *>(p) = v; return; }
    if (nbytes == 4 && (reinterpret_cast<uintptr_t>(p) & 3) == 0) { *reinterpret_cast<uint32_t *>(p) = (uint32_t)v; return; }
    for (uint32_t b = 0; b < nbytes; ++b) p[b] = (uint8_t)(v >> (8 * b));
}

__device__ __forceinline__ uint64_t wire_get(const uint8_t *p, uint32_t nbytes)
{
    if (nbytes == 8 && (reinterpret_cast<uintptr_t>(p) & 7) == 0) return *reinterpret_cast<const uint64_t *>(p);
    if (nbytes == 4 && (reinterpret_cast<uintptr_t>(p) & 3) == 0) return *reinterpret_cast<const uint32_t *>(p);
    uint64_t v = 0;
    for (uint32_t b = 0; b < nbytes; ++b) v |= (uint64_t)p[b] << (8 * b);
    return v;
}

// One warp per item.  WRITE = false: only the item's size is produced (pass 1 of a trimmed pack).
template <bool WRITE>
__global__ void __launch_bounds__(256) rzk_wire_pack_kernel(const __grid_constant__ WireLaunch K)
{
    const uint32_t lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < K.n_items; item += warps) {
        uint8_t *out = WRITE ? K.bytes + K.offsets[item] : nullptr;
        uint64_t pos = 0;
        for (uint32_t q = 0; q < K.ntoks; ++q) {
            const rzk_wire_tok t = K.toks[q];
            if (t.kind == RZK_WIRE_LEN) {
                if (WRITE && lane == 0) wire_put(out + pos, t.value, 8);
                pos += 8;
            } else if (t.kind == RZK_WIRE_TAG) {
                if (WRITE && lane == 0) out[pos] = (uint8_t)t.value;
                pos += 1;
            } else if (t.kind == RZK_WIRE_POLY) {
                const uint64_t poly = (uint64_t)item * K.polys[t.stream] + t.poly;
                const uint32_t len = wire_poly_len(K, t.stream, poly, lane);
                if (WRITE) {
                    if (lane == 0) wire_put(out + pos, len, 8);
                    for (uint32_t i = lane; i < len; i += 32)
                        wire_put(out + pos + 8 + (uint64_t)i * K.elem_bytes, (uint64_t)(int64_t)wire_coeff(K, t.stream, poly, i), K.elem_bytes);
                }
                pos += 8 + (uint64_t)len * K.elem_bytes;
            }
        }
        if (!WRITE && lane == 0) K.sizes[item] = pos;
    }
}

// exclusive prefix sum of n 64-bit sizes into out[0..n] (out[n] = total); one block, a chunk of 1024 per step
__global__ void __launch_bounds__(1024) rzk_wire_scan_kernel(size_t n, const uint64_t *__restrict__ in, uint64_t *__restrict__ out)
{
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry_s;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (size_t base = 0; base < n; base += 1024) {
        const size_t i = base + threadIdx.x;
        const uint64_t v = i < n ? in[i] : 0;
        uint64_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t o = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += o; }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        if (w == 0) {
            uint64_t s = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint64_t o = __shfl_up_sync(0xffffffffu, s, d); if (lane >= (uint32_t)d) s += o; }
            warp_sum[lane] = s;
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        const uint64_t before = carry + (w ? warp_sum[w - 1] : 0) + x - v;
        if (i < n) out[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

// Parses one item per warp.  A literal that differs, a polynomial longer than N, a coefficient that does not fit the
// stream's type, bytes missing or left over: FLAG_FAIL for the item (its outputs are then unspecified but in bounds).
__global__ void __launch_bounds__(256) rzk_wire_unpack_kernel(const __grid_constant__ WireLaunch K)
{
    const uint32_t lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    const int64_t half = (K.q - 1) / 2;
    for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < K.n_items; item += warps) {
        const uint64_t o0 = K.offsets[item], o1 = K.offsets[item + 1];
        uint32_t bad = (o1 < o0 || o1 > K.in_bytes) ? 1u : 0u;          // offsets that leave the buffer: nothing of the item is read
        const uint8_t *in = K.bytes + (bad ? 0 : o0);
        const uint64_t size = bad ? 0 : o1 - o0;
        uint64_t pos = 0;
        for (uint32_t q = 0; q < K.ntoks && !bad; ++q) {
            const rzk_wire_tok t = K.toks[q];
            if (t.kind == RZK_WIRE_LEN) {
                if (pos + 8 > size || wire_get(in + pos, 8) != (uint64_t)t.value) bad = 1;
                pos += 8;
            } else if (t.kind == RZK_WIRE_TAG) {
                if (pos + 1 > size || in[pos] != (uint8_t)t.value) bad = 1;
                pos += 1;
            } else if (t.kind == RZK_WIRE_POLY) {
                uint64_t len = 0;
                if (pos + 8 > size) bad = 1;
                else len = wire_get(in + pos, 8);
                if (len > (uint64_t)kN || pos + 8 + len * K.elem_bytes > size) { bad = 1; len = 0; }
                const uint64_t poly = (uint64_t)item * K.polys[t.stream] + t.poly;
                uint32_t lane_bad = 0;
                for (uint32_t i = lane; i < (uint32_t)kN; i += 32) {
                    int64_t v = 0;
                    if (i < len) {
                        const uint64_t raw = wire_get(in + pos + 8 + (uint64_t)i * K.elem_bytes, K.elem_bytes);
                        v = K.elem_bytes == 8 ? (int64_t)raw : (int64_t)(int32_t)(uint32_t)raw;
                        v %= K.q;                                   // ZqI64::from: any representative -> canonical centred
                        if (v > half) v -= K.q; else if (v < -half) v += K.q;
                    }
                    if (K.dtype[t.stream] == DT_I8) {
                        if (v < -128 || v > 127) lane_bad = 1;
                        reinterpret_cast<int8_t *>(const_cast<void *>(K.base[t.stream]))[poly * kN + i] = (int8_t)v;
                    } else {
                        reinterpret_cast<int32_t *>(const_cast<void *>(K.base[t.stream]))[poly * kN + i] = (int32_t)v;
                    }
                }
                if (__any_sync(0xffffffffu, lane_bad)) bad = 1;
                pos += 8 + len * K.elem_bytes;
            }
        }
        if (!bad && pos != size) bad = 1;
        if (bad && lane == 0) atomicOr(&K.flags[item], FLAG_FAIL);
    }
}

struct WireTokList {
    rzk_wire_tok *t; size_t cap, n; bool overflow;
    void add(uint32_t kind, uint32_t stream, uint32_t poly, uint32_t value)
    {
        if (n < cap && t) { t[n].kind = kind; t[n].stream = stream; t[n].poly = poly; t[n].value = value; }
        else overflow = true;
        ++n;
    }
    void len(uint32_t v) { add(RZK_WIRE_LEN, 0, 0, v); }
    void poly(uint32_t s, uint32_t p) { add(RZK_WIRE_POLY, s, p, 0); }
    void vec(uint32_t s, uint32_t first, uint32_t count) { len(count); for (uint32_t i = 0; i < count; ++i) poly(s, first + i); }   // Vec<Polynomial>
    void mat(uint32_t s, uint32_t first, uint32_t rows) { len(rows); for (uint32_t i = 0; i < rows; ++i) { len(1); poly(s, first + i); } }   // Mat, rows x 1
};

int wire_fill(rzk_engine *e, WireLaunch &K, size_t B, const rzk_wire_tok *toks, size_t ntoks, const rzk_wire_stream *streams, int nstreams,
              int elem_bytes, int trim, rzk_wire_tok **d_toks)
{
    if (!toks || !streams || ntoks == 0 || nstreams < 1 || nstreams > kWireMaxStreams) return fail(e, RZK_ERR_INVALID, "wire: bad token list / stream count");
    if (elem_bytes != 4 && elem_bytes != 8) return fail(e, RZK_ERR_INVALID, "wire: a coefficient is 4 (i32) or 8 (i64, ZqI64) bytes wide");
    if (B >= (1u << 28)) return fail(e, RZK_ERR_INVALID, "wire: more than 2^28 items");
    memset(&K, 0, sizeof(K));
    for (size_t i = 0; i < ntoks; ++i) {
        const rzk_wire_tok &t = toks[i];
        if (t.kind == RZK_WIRE_POLY && ((int)t.stream >= nstreams || !streams[t.stream].base || t.poly >= streams[t.stream].polys_per_item))
            return fail(e, RZK_ERR_INVALID, "wire: a polynomial token names a stream / polynomial that was not supplied");
        if (t.kind > RZK_WIRE_TAG) return fail(e, RZK_ERR_INVALID, "wire: unknown token kind");
    }
    for (int s = 0; s < nstreams; ++s) {
        if (streams[s].dtype > DT_I8) return fail(e, RZK_ERR_INVALID, "wire: stream dtype is 0 (int32) or 1 (int8)");
        K.base[s] = streams[s].base; K.polys[s] = streams[s].polys_per_item; K.dtype[s] = streams[s].dtype;
    }
    if (ntoks * sizeof(rzk_wire_tok) > e->wire_toks_cap) {            // token list staged in a buffer the engine keeps
        RZK_CUDA(e, cudaDeviceSynchronize());
        if (e->d_wire_toks) cudaFree(e->d_wire_toks);
        e->d_wire_toks = nullptr; e->wire_toks_cap = 0;
        const size_t cap = std::max<size_t>(ntoks * sizeof(rzk_wire_tok), 16384);
        RZK_CUDA(e, cudaMalloc(&e->d_wire_toks, cap));
        e->wire_toks_cap = cap;
    }
    *d_toks = reinterpret_cast<rzk_wire_tok *>(e->d_wire_toks);
    K.toks = *d_toks; K.ntoks = (uint32_t)ntoks; K.n_items = (uint32_t)B;
    K.elem_bytes = (uint32_t)elem_bytes; K.trim = trim ? 1u : 0u; K.q = e->P.q;
    return RZK_OK;
}

}  // namespace

extern "C" {

int rzk_wire_layout(int message_kind, uint32_t T, rzk_wire_tok *toks, size_t cap, size_t *ntoks)
{
    if (!ntoks) return RZK_ERR_INVALID;
    WireTokList L{toks, toks ? cap : 0, 0, false};
    const bool needs_T = message_kind == RZK_MSG_SUM_COMMITMENT || message_kind == RZK_MSG_SUM_RESPONSE;
    if (needs_T && (T == 0 || T > 65535)) return RZK_ERR_INVALID;
    switch (message_kind) {
    case RZK_MSG_COMMITMENT:            // Commitment { c: Mat }                                     commit.rs:134-141
        L.mat(0, 0, 2); break;
    case RZK_MSG_OPENING:               // Opening { x: Vec<Polynomial>, r: Mat, f: None }            commit.rs:222-235
        L.vec(0, 0, 1); L.mat(1, 0, 3); L.add(RZK_WIRE_TAG, 0, 0, 0); break;
    case RZK_MSG_OPENING_F:             // ... f: Some(f)
        L.vec(0, 0, 1); L.mat(1, 0, 3); L.add(RZK_WIRE_TAG, 0, 0, 1); L.poly(2, 0); break;
    case RZK_MSG_OPEN_COMMITMENT:       // OpenProofCommitment { c: Commitment, t: Vec<Polynomial> }  open.rs:190-198
        L.mat(0, 0, 2); L.vec(1, 0, 1); break;
    case RZK_MSG_CHALLENGE:             // Open / Linear / Sum ProofChallenge { d: Polynomial }        open.rs:213-219, linear.rs:309-315, sum.rs:375-381
        L.poly(0, 0); break;
    case RZK_MSG_OPEN_RESPONSE:         // OpenProofResponse { z: Mat }                               open.rs:222-228
        L.mat(0, 0, 3); break;
    case RZK_MSG_LINEAR_COMMITMENT:     // LinearProofCommitment { c, cp, g, t, tp, u: Mat }          linear.rs:271-285
        L.mat(0, 0, 2); L.mat(1, 0, 2); L.poly(2, 0); L.vec(3, 0, 1); L.vec(4, 0, 1); L.mat(5, 0, 1); break;
    case RZK_MSG_LINEAR_RESPONSE:       // LinearProofResponse { z: Mat, zp: Mat } -- the reference does not derive Serialize for it
        L.mat(0, 0, 3); L.mat(1, 0, 3); break;                                        // (linear.rs:318): the layout the derive WOULD give
    case RZK_MSG_SUM_COMMITMENT:        // SumProofCommitment { cp, cs: Vec<Commitment>, gs, tp, ts: Vec<Vec<Polynomial>>, u }   sum.rs:342-355
        L.mat(0, 0, 2);
        L.len(T); for (uint32_t i = 0; i < T; ++i) L.mat(1, 2 * i, 2);
        L.vec(2, 0, T); L.vec(3, 0, 1);
        L.len(T); for (uint32_t i = 0; i < T; ++i) L.vec(4, i, 1);
        L.mat(5, 0, 1); break;
    case RZK_MSG_SUM_RESPONSE:          // SumProofResponse { zp: Mat, zs: Vec<Mat> }                 sum.rs:384-391
        L.mat(0, 0, 3);
        L.len(T); for (uint32_t i = 0; i < T; ++i) L.mat(1, 3 * i, 3);
        break;
    default:
        return RZK_ERR_INVALID;
    }
    *ntoks = L.n;
    return (toks && L.overflow) ? RZK_ERR_INVALID : RZK_OK;
}

int rzk_wire_pack_dev(rzk_engine *e, size_t B, const rzk_wire_tok *toks, size_t ntoks, const rzk_wire_stream *streams, int nstreams,
                      int elem_bytes, int trim, uint8_t *out, size_t out_capacity, uint64_t *offsets, uint64_t *total_bytes, void *stream)
{
    RZK_TRY(check_ready(e, false));
    if (!offsets || !total_bytes) return fail(e, RZK_ERR_INVALID, "wire: null argument");
    Guard g(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    WireLaunch K;
    rzk_wire_tok *d_toks = nullptr;
    RZK_TRY(wire_fill(e, K, B, toks, ntoks, streams, nstreams, elem_bytes, trim, &d_toks));
    auto body = [&]() -> int {
        RZK_CUDA(e, cudaMemcpyAsync(d_toks, toks, ntoks * sizeof(rzk_wire_tok), cudaMemcpyHostToDevice, s));
        const unsigned grid = (unsigned)std::min<size_t>((B + 7) / 8 + 1, (size_t)e->num_sms * 8);
        // pass 1: sizes (into the offsets array, shifted by one word so that the scan may run in place is NOT assumed: scratch)
        RZK_TRY(ensure_scratch(e, (B + 1) * sizeof(uint64_t)));
        K.sizes = reinterpret_cast<uint64_t *>(e->scratch);
        if (B) { rzk_wire_pack_kernel<false><<<grid, 256, 0, s>>>(K); RZK_CUDA(e, cudaGetLastError()); e->launches++; }
        rzk_wire_scan_kernel<<<1, 1024, 0, s>>>(B, K.sizes, offsets);
        RZK_CUDA(e, cudaGetLastError());
        e->launches++;
        RZK_CUDA(e, cudaMemcpyAsync(total_bytes, offsets + B, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        RZK_CUDA(e, cudaStreamSynchronize(s));
        if (!out) return RZK_OK;                              // size query
        if (*total_bytes > out_capacity) return fail(e, RZK_ERR_INVALID, "wire: output buffer too small (total_bytes holds the size needed)");
        K.bytes = out; K.offsets = offsets;
        if (B) { rzk_wire_pack_kernel<true><<<grid, 256, 0, s>>>(K); RZK_CUDA(e, cudaGetLastError()); e->launches++; }
        RZK_CUDA(e, cudaStreamSynchronize(s));
        return RZK_OK;
    };
    return body();
}

int rzk_wire_unpack_dev(rzk_engine *e, size_t B, const rzk_wire_tok *toks, size_t ntoks, const rzk_wire_stream *streams, int nstreams,
                        int elem_bytes, const uint8_t *in, size_t in_bytes, const uint64_t *offsets, uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e, false));
    if (!in || !offsets || !flags) return fail(e, RZK_ERR_INVALID, "wire: null argument");
    Guard g(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    WireLaunch K;
    rzk_wire_tok *d_toks = nullptr;
    RZK_TRY(wire_fill(e, K, B, toks, ntoks, streams, nstreams, elem_bytes, 0, &d_toks));
    auto body = [&]() -> int {
        RZK_CUDA(e, cudaMemcpyAsync(d_toks, toks, ntoks * sizeof(rzk_wire_tok), cudaMemcpyHostToDevice, s));
        K.bytes = const_cast<uint8_t *>(in); K.offsets = offsets; K.flags = flags; K.in_bytes = in_bytes;
        const unsigned grid = (unsigned)std::min<size_t>((B + 7) / 8 + 1, (size_t)e->num_sms * 8);
        if (B) { rzk_wire_unpack_kernel<<<grid, 256, 0, s>>>(K); RZK_CUDA(e, cudaGetLastError()); e->launches++; }
        RZK_CUDA(e, cudaStreamSynchronize(s));
        return RZK_OK;
    };
    return body();
}

// ---- host forms: the streams, the bytes and the offsets are host arrays (staged through device buffers of the call) ----

namespace {
struct WireStage {
    std::vector<void *> dev;
    ~WireStage() { for (void *p : dev) if (p) cudaFree(p); }
    int alloc(rzk_engine *e, void **out, size_t bytes)
    {
        *out = nullptr;
        RZK_CUDA(e, cudaMalloc(out, std::max<size_t>(bytes, 16)));
        dev.push_back(*out);
        return RZK_OK;
    }
};
size_t wire_stream_bytes(const rzk_wire_stream &s, size_t B) { return B * s.polys_per_item * (size_t)kN * (s.dtype == DT_I8 ? 1 : 4); }
}  // namespace

int rzk_wire_pack(rzk_engine *e, size_t B, const rzk_wire_tok *toks, size_t ntoks, const rzk_wire_stream *streams, int nstreams,
                  int elem_bytes, int trim, uint8_t *out, size_t out_capacity, uint64_t *offsets, uint64_t *total_bytes)
{
    RZK_TRY(check_ready(e, false));
    if (!streams || !offsets || !total_bytes || nstreams < 1 || nstreams > kWireMaxStreams) return fail(e, RZK_ERR_INVALID, "wire: bad argument");
    Guard g(e->device);
    WireStage st;
    rzk_wire_stream ds[kWireMaxStreams];
    for (int i = 0; i < nstreams; ++i) {
        ds[i] = streams[i];
        if (!streams[i].base) continue;
        void *p;
        RZK_TRY(st.alloc(e, &p, wire_stream_bytes(streams[i], B)));
        RZK_CUDA(e, cudaMemcpy(p, streams[i].base, wire_stream_bytes(streams[i], B), cudaMemcpyHostToDevice));
        ds[i].base = p;
    }
    void *d_off, *d_out = nullptr;
    RZK_TRY(st.alloc(e, &d_off, (B + 1) * sizeof(uint64_t)));
    RZK_TRY(rzk_wire_pack_dev(e, B, toks, ntoks, ds, nstreams, elem_bytes, trim, nullptr, 0, (uint64_t *)d_off, total_bytes, nullptr));
    RZK_CUDA(e, cudaMemcpy(offsets, d_off, (B + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (!out) return RZK_OK;                                  // size query
    if (*total_bytes > out_capacity) return fail(e, RZK_ERR_INVALID, "wire: output buffer too small (total_bytes holds the size needed)");
    RZK_TRY(st.alloc(e, &d_out, *total_bytes));
    RZK_TRY(rzk_wire_pack_dev(e, B, toks, ntoks, ds, nstreams, elem_bytes, trim, (uint8_t *)d_out, *total_bytes, (uint64_t *)d_off, total_bytes, nullptr));
    RZK_CUDA(e, cudaMemcpy(out, d_out, *total_bytes, cudaMemcpyDeviceToHost));
    return RZK_OK;
}

int rzk_wire_unpack(rzk_engine *e, size_t B, const rzk_wire_tok *toks, size_t ntoks, const rzk_wire_stream *streams, int nstreams,
                    int elem_bytes, const uint8_t *in, size_t in_bytes, const uint64_t *offsets, uint8_t *ok_bitmap)
{
    RZK_TRY(check_ready(e, false));
    if (!streams || !in || !offsets || !ok_bitmap || nstreams < 1 || nstreams > kWireMaxStreams) return fail(e, RZK_ERR_INVALID, "wire: bad argument");
    Guard g(e->device);
    WireStage st;
    rzk_wire_stream ds[kWireMaxStreams];
    for (int i = 0; i < nstreams; ++i) {
        ds[i] = streams[i];
        if (!streams[i].base) continue;
        void *p;
        RZK_TRY(st.alloc(e, &p, wire_stream_bytes(streams[i], B)));
        RZK_CUDA(e, cudaMemset(p, 0, wire_stream_bytes(streams[i], B)));
        ds[i].base = p;
    }
    void *d_in, *d_off, *d_flags, *d_bm;
    RZK_TRY(st.alloc(e, &d_in, in_bytes));
    RZK_TRY(st.alloc(e, &d_off, (B + 1) * sizeof(uint64_t)));
    RZK_TRY(st.alloc(e, &d_flags, (B + 1) * sizeof(uint32_t)));
    RZK_TRY(st.alloc(e, &d_bm, (B + 7) / 8 + 1));
    RZK_CUDA(e, cudaMemcpy(d_in, in, in_bytes, cudaMemcpyHostToDevice));
    RZK_CUDA(e, cudaMemcpy(d_off, offsets, (B + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    RZK_CUDA(e, cudaMemset(d_flags, 0, (B + 1) * sizeof(uint32_t)));
    RZK_TRY(rzk_wire_unpack_dev(e, B, toks, ntoks, ds, nstreams, elem_bytes, (const uint8_t *)d_in, in_bytes, (const uint64_t *)d_off,
                                (uint32_t *)d_flags, nullptr));
    RZK_TRY(rzk_flags_to_bitmap_dev(e, B, (const uint32_t *)d_flags, (uint8_t *)d_bm, nullptr, nullptr));
    RZK_CUDA(e, cudaMemcpy(ok_bitmap, d_bm, (B + 7) / 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nstreams; ++i)
        if (streams[i].base)
            RZK_CUDA(e, cudaMemcpy(const_cast<void *>(streams[i].base), ds[i].base, wire_stream_bytes(streams[i], B), cudaMemcpyDeviceToHost));
    return RZK_OK;
}

}  // extern "C"
