// rzk_fs.cuh -- Fiat-Shamir challenges on the device.
//
// The reference is interactive: the verifier draws d with its own RNG (open.rs:143-158, challenge_space.rs:12-33); its README
// only names the transform as a possibility (README.md:16).  This file implements the transcript specified in
// docs/FIAT_SHAMIR.md so that a batch of proofs needs no round trip to the verifier -- or to the host:
//
//     d_i = SampleInBall_kappa( SHAKE128( prefix || poly_0 || poly_1 || ... ) )        one hash per batch item i
//
// where every polynomial of item i's first message is absorbed as N little-endian int32 canonical centred coefficients, in the
// field order of the reference's commitment struct (open.rs:190-198, linear.rs:271-285, sum.rs:342-355), and `prefix` (a
// multiple of 8 bytes: domain tag, key digest, shape words) is supplied by the caller.  SampleInBall is the inside-out
// Fisher-Yates walk of CRYSTALS-Dilithium adapted to N = 512: the first 8 squeezed bytes are the sign bits, then for
// i = N - kappa .. N - 1 a position j <= i is drawn by rejection from 16-bit words masked to 9 bits, d[i] = d[j], d[j] = +-1:
// exactly kappa entries +-1, uniform over the challenge space of challenge_space.rs:12-33 given a uniform hash output.
// One thread per item: the 25-word Keccak state lives in registers; the 4 to 26 KB of an item are read with 16-byte loads.
// oracle/fs_ref.py restates the same function with hashlib.shake_128 (checker only).
#pragma once

namespace {

constexpr int kFsMaxSegs = 8;

struct FsLaunch {
    const uint64_t *prefix;       // device copy, prefix_lanes 64-bit words
    uint32_t prefix_lanes;
    uint32_t n_items, kappa, nsegs;
    const void *base[kFsMaxSegs];
    uint32_t polys[kFsMaxSegs];
    uint32_t dtype[kFsMaxSegs];
    uint32_t div[kFsMaxSegs];     // item group = item / div (a per-instance segment shared by the T terms of an instance)
    int8_t *d;                    // [n_items][N]
};

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int s) { return (x << s) | (x >> (64 - s)); }

__constant__ uint64_t kKeccakRC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
        0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
        0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
        0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};

__device__ void keccak_f1600(uint64_t (&a)[25])
{
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20], c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21];
        uint64_t c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22], c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23];
        uint64_t c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
        const uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
#pragma unroll
        for (int y = 0; y < 25; y += 5) { a[y] ^= d0; a[y + 1] ^= d1; a[y + 2] ^= d2; a[y + 3] ^= d3; a[y + 4] ^= d4; }
        // rho + pi
        uint64_t b[25];
        b[0] = a[0];
        b[10] = rotl64(a[1], 1);   b[20] = rotl64(a[2], 62);  b[5] = rotl64(a[3], 28);   b[15] = rotl64(a[4], 27);
        b[16] = rotl64(a[5], 36);  b[1] = rotl64(a[6], 44);   b[11] = rotl64(a[7], 6);   b[21] = rotl64(a[8], 55);  b[6] = rotl64(a[9], 20);
        b[7] = rotl64(a[10], 3);   b[17] = rotl64(a[11], 10); b[2] = rotl64(a[12], 43);  b[12] = rotl64(a[13], 25); b[22] = rotl64(a[14], 39);
        b[23] = rotl64(a[15], 41); b[8] = rotl64(a[16], 45);  b[18] = rotl64(a[17], 15); b[3] = rotl64(a[18], 21);  b[13] = rotl64(a[19], 8);
        b[14] = rotl64(a[20], 18); b[24] = rotl64(a[21], 2);  b[9] = rotl64(a[22], 61);  b[19] = rotl64(a[23], 56); b[4] = rotl64(a[24], 14);
        // chi
#pragma unroll
        for (int y = 0; y < 25; y += 5) {
            a[y] = b[y] ^ (~b[y + 1] & b[y + 2]);
            a[y + 1] = b[y + 1] ^ (~b[y + 2] & b[y + 3]);
            a[y + 2] = b[y + 2] ^ (~b[y + 3] & b[y + 4]);
            a[y + 3] = b[y + 3] ^ (~b[y + 4] & b[y]);
            a[y + 4] = b[y + 4] ^ (~b[y] & b[y + 1]);
        }
        a[0] ^= kKeccakRC[r];
    }
}

struct Shake128 {
    uint64_t s[25];
    int pos;                        // next lane of the 21-lane (168-byte) rate
    __device__ void init()
    {
#pragma unroll
        for (int i = 0; i < 25; ++i) s[i] = 0;
        pos = 0;
    }
    __device__ __forceinline__ void absorb(uint64_t v)
    {
        // (the state is indexed dynamically only here; the permutation works on registers)
        switch (pos) {
#define RZK_FS_CASE(i) case i: s[i] ^= v; break;
            RZK_FS_CASE(0) RZK_FS_CASE(1) RZK_FS_CASE(2) RZK_FS_CASE(3) RZK_FS_CASE(4) RZK_FS_CASE(5) RZK_FS_CASE(6)
            RZK_FS_CASE(7) RZK_FS_CASE(8) RZK_FS_CASE(9) RZK_FS_CASE(10) RZK_FS_CASE(11) RZK_FS_CASE(12) RZK_FS_CASE(13)
            RZK_FS_CASE(14) RZK_FS_CASE(15) RZK_FS_CASE(16) RZK_FS_CASE(17) RZK_FS_CASE(18) RZK_FS_CASE(19) RZK_FS_CASE(20)
        }
        if (++pos == 21) { keccak_f1600(s); pos = 0; }
    }
    __device__ void finish()        // SHAKE domain bits 1111 + pad10*1; the message is a whole number of lanes
    {
        absorb_pad();
        keccak_f1600(s);
        pos = 0;
    }
    __device__ __forceinline__ void absorb_pad()
    {
        const uint64_t v = 0x1Full;
        switch (pos) {
            RZK_FS_CASE(0) RZK_FS_CASE(1) RZK_FS_CASE(2) RZK_FS_CASE(3) RZK_FS_CASE(4) RZK_FS_CASE(5) RZK_FS_CASE(6)
            RZK_FS_CASE(7) RZK_FS_CASE(8) RZK_FS_CASE(9) RZK_FS_CASE(10) RZK_FS_CASE(11) RZK_FS_CASE(12) RZK_FS_CASE(13)
            RZK_FS_CASE(14) RZK_FS_CASE(15) RZK_FS_CASE(16) RZK_FS_CASE(17) RZK_FS_CASE(18) RZK_FS_CASE(19) RZK_FS_CASE(20)
#undef RZK_FS_CASE
        }
        s[20] ^= 0x8000000000000000ull;
    }
    __device__ __forceinline__ uint64_t squeeze()
    {
        if (pos == 21) { keccak_f1600(s); pos = 0; }
        uint64_t v = 0;
        switch (pos) {
#define RZK_FS_CASE(i) case i: v = s[i]; break;
            RZK_FS_CASE(0) RZK_FS_CASE(1) RZK_FS_CASE(2) RZK_FS_CASE(3) RZK_FS_CASE(4) RZK_FS_CASE(5) RZK_FS_CASE(6)
            RZK_FS_CASE(7) RZK_FS_CASE(8) RZK_FS_CASE(9) RZK_FS_CASE(10) RZK_FS_CASE(11) RZK_FS_CASE(12) RZK_FS_CASE(13)
            RZK_FS_CASE(14) RZK_FS_CASE(15) RZK_FS_CASE(16) RZK_FS_CASE(17) RZK_FS_CASE(18) RZK_FS_CASE(19) RZK_FS_CASE(20)
#undef RZK_FS_CASE
        }
        ++pos;
        return v;
    }
};

__global__ void __launch_bounds__(128) rzk_fs_challenge_kernel(const __grid_constant__ FsLaunch K)
{
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= K.n_items) return;
    Shake128 H;
    H.init();
    for (uint32_t i = 0; i < K.prefix_lanes; ++i) H.absorb(K.prefix[i]);
    for (uint32_t sg = 0; sg < K.nsegs; ++sg) {
        const uint64_t first = (uint64_t)(item / K.div[sg]) * K.polys[sg];
        const uint32_t words = K.polys[sg] * (uint32_t)kN;                       // coefficients of this segment
        if (K.dtype[sg] == DT_I8) {
            const int8_t *src = reinterpret_cast<const int8_t *>(K.base[sg]) + first * kN;
            for (uint32_t i = 0; i < words; i += 8) {
                const int2 q = __ldg(reinterpret_cast<const int2 *>(src + i));
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int w = h < 2 ? q.x : q.y;
                    const int32_t c0 = (int8_t)(w >> (16 * (h & 1))), c1 = (int8_t)(w >> (16 * (h & 1) + 8));
                    H.absorb((uint64_t)(uint32_t)c0 | ((uint64_t)(uint32_t)c1 << 32));
                }
            }
        } else {
            const int32_t *src = reinterpret_cast<const int32_t *>(K.base[sg]) + first * kN;
            for (uint32_t i = 0; i < words; i += 4) {
                const int4 q = __ldg(reinterpret_cast<const int4 *>(src + i));
                H.absorb((uint64_t)(uint32_t)q.x | ((uint64_t)(uint32_t)q.y << 32));
                H.absorb((uint64_t)(uint32_t)q.z | ((uint64_t)(uint32_t)q.w << 32));
            }
        }
    }
    H.finish();
    // SampleInBall: kappa entries +-1 (kappa <= 64 sign bits)
    int8_t *d = K.d + (size_t)item * kN;
    for (int i = 0; i < kN / 16; ++i) reinterpret_cast<uint4 *>(d)[i] = make_uint4(0, 0, 0, 0);
    uint64_t signs = H.squeeze();
    uint64_t buf = 0;
    int have = 0;                                                               // 16-bit words left in buf
    for (uint32_t i = (uint32_t)kN - K.kappa; i < (uint32_t)kN; ++i) {
        uint32_t j;
        do {
            if (have == 0) { buf = H.squeeze(); have = 4; }
            j = (uint32_t)buf & 0x1FFu;
            buf >>= 16; --have;
        } while (j > i);
        d[i] = d[j];
        d[j] = (signs & 1) ? (int8_t)-1 : (int8_t)1;
        signs >>= 1;
    }
}

int dev_fs_challenge(rzk_engine *e, size_t B, const uint64_t *d_prefix, uint32_t prefix_lanes, const rzk_wire_stream *segs,
                     const uint32_t *divs, int nsegs, int8_t *d, cudaStream_t s)
{
    if (B == 0) return RZK_OK;
    if (nsegs < 1 || nsegs > kFsMaxSegs) return fail(e, RZK_ERR_INVALID, "fs: 1 to 8 transcript segments");
    if (e->P.kappa > 64) return fail(e, RZK_ERR_UNSUPPORTED, "fs: SampleInBall carries 64 sign bits (kappa <= 64)");
    FsLaunch K;
    memset(&K, 0, sizeof(K));
    K.prefix = d_prefix; K.prefix_lanes = prefix_lanes; K.n_items = (uint32_t)B; K.kappa = (uint32_t)std::min<int64_t>(e->P.kappa, kN);
    K.nsegs = (uint32_t)nsegs; K.d = d;
    for (int i = 0; i < nsegs; ++i) {
        if (!segs[i].base || segs[i].polys_per_item == 0 || segs[i].dtype > DT_I8) return fail(e, RZK_ERR_INVALID, "fs: bad transcript segment");
        K.base[i] = segs[i].base; K.polys[i] = segs[i].polys_per_item; K.dtype[i] = segs[i].dtype; K.div[i] = divs ? divs[i] : 1u;
        if (K.div[i] == 0) return fail(e, RZK_ERR_INVALID, "fs: segment group size 0");
    }
    rzk_fs_challenge_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(K);
    RZK_CUDA(e, cudaGetLastError());
    e->launches++;
    return RZK_OK;
}

// the caller's prefix, staged on the device (kept by the engine, grown on demand)
int fs_stage_prefix(rzk_engine *e, const uint8_t *prefix, size_t prefix_len, cudaStream_t s)
{
    if (prefix_len % 8 != 0) return fail(e, RZK_ERR_INVALID, "fs: the transcript prefix is a multiple of 8 bytes (pad the domain tag with zeros)");
    if (prefix_len && !prefix) return fail(e, RZK_ERR_INVALID, "fs: null prefix");
    if (prefix_len > e->fs_prefix_cap) {
        RZK_CUDA(e, cudaDeviceSynchronize());
        if (e->d_fs_prefix) cudaFree(e->d_fs_prefix);
        e->d_fs_prefix = nullptr; e->fs_prefix_cap = 0;
        RZK_CUDA(e, cudaMalloc(&e->d_fs_prefix, std::max<size_t>(prefix_len, 256)));
        e->fs_prefix_cap = std::max<size_t>(prefix_len, 256);
    }
    // (pageable source: the copy is staged by the runtime before the call returns)
    if (prefix_len) RZK_CUDA(e, cudaMemcpyAsync(e->d_fs_prefix, prefix, prefix_len, cudaMemcpyHostToDevice, s));
    return RZK_OK;
}

}  // namespace

extern "C" {

int rzk_fs_challenge_dev(rzk_engine *e, size_t B, const uint8_t *prefix, size_t prefix_len, const rzk_wire_stream *segs, int nsegs,
                         int8_t *d, void *stream)
{
    RZK_TRY(check_ready(e, false));
    if (!segs || !d) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    RZK_TRY(fs_stage_prefix(e, prefix, prefix_len, (cudaStream_t)stream));
    return dev_fs_challenge(e, B, reinterpret_cast<const uint64_t *>(e->d_fs_prefix), (uint32_t)(prefix_len / 8), segs, nullptr, nsegs, d,
                            (cudaStream_t)stream);
}

int rzk_open_prove_fs_batch_dev(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, const int32_t *y, const uint8_t *prefix,
                                size_t prefix_len, int32_t *c, int32_t *t, int8_t *d, int32_t *z, uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({x, r, y, c, t, d, z, flags})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    RZK_TRY(fs_stage_prefix(e, prefix, prefix_len, s));
    RZK_TRY(ensure_need(e, B + 1));
    RZK_TRY(dev_open_commit(e, B, x, r, y, c, t, flags, s));                                   // open.rs:80-103
    const rzk_wire_stream segs[2] = {{c, 2, DT_I32}, {t, 1, DT_I32}};                          // OpenProofCommitment { c, t }
    RZK_TRY(dev_fs_challenge(e, B, reinterpret_cast<const uint64_t *>(e->d_fs_prefix), (uint32_t)(prefix_len / 8), segs, nullptr, 2, d, s));
    return dev_respond(e, B, y, r, d, 1, z, e->d_need, s);                                      // open.rs:107-117
}

int rzk_open_verify_fs_batch_dev(rzk_engine *e, size_t B, const int32_t *c, const int32_t *t, const int32_t *z, const uint8_t *prefix,
                                 size_t prefix_len, int8_t *d, uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({c, t, z, d, flags})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    RZK_TRY(fs_stage_prefix(e, prefix, prefix_len, s));
    const rzk_wire_stream segs[2] = {{c, 2, DT_I32}, {t, 1, DT_I32}};
    RZK_TRY(dev_fs_challenge(e, B, reinterpret_cast<const uint64_t *>(e->d_fs_prefix), (uint32_t)(prefix_len / 8), segs, nullptr, 2, d, s));
    return dev_verify_first(e, B, z, t, c, 2, d, 1, nullptr, flags, 1, s);                      // open.rs:162-174
}

// ---- host entry points (host pointers, chunked pipeline as the other host calls) ----

int rzk_open_prove_fs_batch(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, const int32_t *y, const uint8_t *prefix,
                            size_t prefix_len, int32_t *c, int32_t *t, int8_t *d, int32_t *z, uint8_t *ok)
{
    RZK_TRY(check_ready(e));
    if (any_null({x, r, y, c, t, d, z, ok})) return fail(e, RZK_ERR_INVALID, "null argument");
    {
        Guard g(e->device);
        RZK_TRY(fs_stage_prefix(e, prefix, prefix_len, e->pipe[0].stream));       // once; every pipeline stream reads it
        RZK_CUDA(e, cudaStreamSynchronize(e->pipe[0].stream));
    }
    const uint64_t *pre = reinterpret_cast<const uint64_t *>(e->d_fs_prefix);
    const uint32_t lanes = (uint32_t)(prefix_len / 8);
    std::vector<HArr> a = {{x, nullptr, kPolyBytes}, {r, nullptr, 3 * kN}, {y, nullptr, 3 * kPolyBytes},
                           {nullptr, c, 2 * kPolyBytes}, {nullptr, t, kPolyBytes}, {nullptr, d, kN}, {nullptr, z, 3 * kPolyBytes}};
    return run_chunked(e, B, a, 2 * sizeof(uint32_t), ok, [&](size_t n, void **p, char *sc, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        RZK_TRY(dev_open_commit(e, n, (const int32_t *)p[0], (const int8_t *)p[1], (const int32_t *)p[2], (int32_t *)p[3], (int32_t *)p[4], fl, s, rm));
        const rzk_wire_stream segs[2] = {{p[3], 2, DT_I32}, {p[4], 1, DT_I32}};
        RZK_TRY(dev_fs_challenge(e, n, pre, lanes, segs, nullptr, 2, (int8_t *)p[5], s));
        return dev_respond(e, n, (const int32_t *)p[2], (const int8_t *)p[1], (const int8_t *)p[5], 1, (int32_t *)p[6], (uint32_t *)sc, s);
    });
}

int rzk_open_verify_fs_batch(rzk_engine *e, size_t B, const int32_t *c, const int32_t *t, const int32_t *z, const uint8_t *prefix,
                             size_t prefix_len, uint8_t *bm)
{
    RZK_TRY(check_ready(e));
    if (any_null({c, t, z, bm})) return fail(e, RZK_ERR_INVALID, "null argument");
    {
        Guard g(e->device);
        RZK_TRY(fs_stage_prefix(e, prefix, prefix_len, e->pipe[0].stream));
        RZK_CUDA(e, cudaStreamSynchronize(e->pipe[0].stream));
    }
    const uint64_t *pre = reinterpret_cast<const uint64_t *>(e->d_fs_prefix);
    const uint32_t lanes = (uint32_t)(prefix_len / 8);
    std::vector<HArr> a = {{c, nullptr, 2 * kPolyBytes}, {t, nullptr, kPolyBytes}, {z, nullptr, 3 * kPolyBytes}};
    return run_chunked(e, B, a, kN, bm, [&](size_t n, void **p, char *sc, uint32_t *fl, cudaStream_t s, uint32_t *) {
        const rzk_wire_stream segs[2] = {{p[0], 2, DT_I32}, {p[1], 1, DT_I32}};
        RZK_TRY(dev_fs_challenge(e, n, pre, lanes, segs, nullptr, 2, (int8_t *)sc, s));
        return dev_verify_first(e, n, (const int32_t *)p[2], (const int32_t *)p[1], (const int32_t *)p[0], 2, (const int8_t *)sc, 1,
                                nullptr, fl, 1, s);
    });
}

}  // extern "C"
