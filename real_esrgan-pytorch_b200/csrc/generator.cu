// RRDBNet x4 generator forward: host orchestration + layout / weight-packing kernels + C ABI.
// Reference: /root/reference/model.py:64-132 (RDB, RRDB), :206-275 (Generator). See DESIGN.md §3-4.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/resr.h"
#include "conv3x3.cuh"
#include "errors.h"
#include "gen_internal.cuh"

namespace resr {

// ------------------------------------------------------------------------------------------- small kernels

// OIHW fp32 -> [slice][chunk][dx][n = dy*NOUT + co][64 ch] 16-bit with the 128B shared-memory swizzle applied, so a
// plain bulk copy drops a ready-to-use UMMA B operand into shared memory.
__device__ __forceinline__ void pack_conv_body(const float* __restrict__ w, const float* __restrict__ bias, uint16_t* __restrict__ wp,
                                               float* __restrict__ bp, int cin, int cout, int nout, int nslices, int nchunks, int fmt,
                                               int transposed, size_t first, size_t stride) {
    const int NT = 3 * nout;
    const size_t total = static_cast<size_t>(nslices) * nchunks * 3 * NT * 64;
    for (size_t idx = first; idx < total; idx += stride) {   // (a layer's pack has < 2^32 elements: 32-bit index decoding)
        const int k = static_cast<int>(idx & 63);
        unsigned t = static_cast<unsigned>(idx >> 6);
        const int n = t % NT;
        t /= NT;
        const int dx = t % 3;
        t /= 3;
        const int chunk = t % nchunks;
        const int slice = static_cast<int>(t / nchunks);
        const int co = slice * nout + n % nout;
        const int dy = n / nout;
        const int ci = chunk * 64 + k;
        float v = 0.f;
        if (!transposed) {
            if (co < cout && ci < cin) v = w[((static_cast<size_t>(co) * cin + ci) * 3 + dy) * 3 + dx];
        } else {
            // data-gradient convolution: its output channel `co` is a forward INPUT channel, its input channel `ci` a
            // forward OUTPUT channel, taps flipped
            if (co < cin && ci < cout) v = w[((static_cast<size_t>(ci) * cin + co) * 3 + (2 - dy)) * 3 + (2 - dx)];
        }
        uint16_t bits;
        if (fmt == 1) {
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            bits = *reinterpret_cast<uint16_t*>(&h);
        } else {
            __half h = __float2half_rn(v);
            bits = *reinterpret_cast<uint16_t*>(&h);
        }
        const size_t tile = ((static_cast<size_t>(slice) * nchunks + chunk) * 3 + dx) * NT * 64;  // elements
        const size_t off = tile + static_cast<size_t>(n) * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7));
        wp[off] = bits;
    }
    const int nb = nslices * nout;
    if (bp)
        for (size_t i = first; i < static_cast<size_t>(nb); i += stride) bp[i] = (i < static_cast<size_t>(cout) && bias) ? bias[i] : 0.f;
}

__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ bias, uint16_t* __restrict__ wp,
                                 float* __restrict__ bp, int cin, int cout, int nout, int nslices, int nchunks, int fmt,
                                 int transposed) {
    pack_conv_body(w, bias, wp, bp, cin, cout, nout, nslices, nchunks, fmt, transposed,
                   blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x, static_cast<size_t>(gridDim.x) * blockDim.x);
}

// All layers in one launch (blockIdx.y = layer): the repack after every optimizer step is 2 launches instead of 701.
struct PackJob {
    unsigned long long p_off, w_off, b_off;  // floats into the flat parameter vector, bytes into the pack, floats into the bias
    int cin, cout, nout, nslices, nchunks, fmt;
};
__global__ void __launch_bounds__(256) pack_all_kernel(const float* __restrict__ flat, const PackJob* __restrict__ jobs,
                                                       uint8_t* __restrict__ wpack, float* __restrict__ bias, int transposed) {
    const PackJob j = jobs[blockIdx.y];
    const float* w = flat + j.p_off;
    const float* b = w + static_cast<size_t>(j.cout) * j.cin * 9;
    pack_conv_body(w, transposed ? nullptr : b, reinterpret_cast<uint16_t*>(wpack + j.w_off), transposed ? nullptr : bias + j.b_off,
                   j.cin, j.cout, j.nout, j.nslices, j.nchunks, j.fmt, transposed,
                   blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x, static_cast<size_t>(gridDim.x) * blockDim.x);
}

// Data-gradient packs of a dense block in "mirrored dense block" form (train.cu backward): step b (4, 3, 2, 1, 0) maps
//     dYcat' = [dY5 (64) | dY4 (32) | dY3 | dY2 | dY1]   (its first 64 + 32 * (4 - b) channels = the layers AFTER block b)
// to the gradient of concat-buffer block b (b = 0: the block input x, 64 channels; b >= 1: out_b, 32 channels):
//     W'_b[co'][ci'][dy][dx] = W_l[co_l][cin_index(b, co')][2 - dy][2 - dx],   (l, co_l) = the layer / channel ci' belongs to.
// Step b has exactly the shape of forward layer conv(5 - b) (Cin 64 / 96 / 128 / 160 / 192, Cout 32 / 32 / 32 / 32 / 64), so
// its tiles are stored at that layer's pack offset in a second pack buffer. blockIdx.y = step, blockIdx.z = dense block.
__global__ void __launch_bounds__(256) pack_rdb_bwd_kernel(const float* __restrict__ flat, const PackJob* __restrict__ jobs,
                                                          uint8_t* __restrict__ wpack) {
    const int r = blockIdx.z, b = 4 - static_cast<int>(blockIdx.y);     // blockIdx.y = 0..4 <-> b = 4..0 <-> forward shape conv1..conv5
    const PackJob shape = jobs[1 + 5 * r + blockIdx.y];                 // pack geometry (offset, nout, nslices, nchunks) of that shape
    const int cin_p = 64 + 32 * (4 - b);                                // channels of dYcat' this step reads
    const int cout_p = b == 0 ? 64 : 32;
    const int NT = 3 * shape.nout;
    uint16_t* wp = reinterpret_cast<uint16_t*>(wpack + shape.w_off);
    const size_t total = static_cast<size_t>(shape.nslices) * shape.nchunks * 3 * NT * 64;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(idx & 63);
        unsigned t = static_cast<unsigned>(idx >> 6);   // (a layer's pack has < 2^32 elements: 32-bit index decoding)
        const int n = t % NT;
        t /= NT;
        const int dx = t % 3;
        t /= 3;
        const int chunk = t % shape.nchunks;
        const int slice = static_cast<int>(t / shape.nchunks);
        const int co = slice * shape.nout + n % shape.nout;   // channel inside concat block b
        const int dy = n / shape.nout;
        const int ci = chunk * 64 + k;                          // channel of dYcat'
        float v = 0.f;
        if (co < cout_p && ci < cin_p) {
            const int l = ci < 64 ? 5 : 4 - (ci - 64) / 32;     // forward layer (1..5) whose output gradient this channel is
            const int co_l = ci < 64 ? ci : (ci - 64) % 32;
            const int cin_l = 64 + 32 * (l - 1);
            const int cin_idx = b == 0 ? co : 64 + 32 * (b - 1) + co;   // input channel of layer l that is block b's channel co
            const PackJob lay = jobs[1 + 5 * r + (l - 1)];
            v = flat[lay.p_off + ((static_cast<size_t>(co_l) * cin_l + cin_idx) * 3 + (2 - dy)) * 3 + (2 - dx)];
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const size_t tile = ((static_cast<size_t>(slice) * shape.nchunks + chunk) * 3 + dx) * NT * 64;  // elements
        wp[tile + static_cast<size_t>(n) * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7))] = *reinterpret_cast<const uint16_t*>(&h);
    }
}

// NCHW fp32 -> NHWC 16-bit with channels zero-padded to c_pad (multiple of 8).
__global__ void nchw_to_nhwc16_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, int N, int C, int H, int W,
                                      int c_pad, int fmt) {
    const size_t npix = static_cast<size_t>(N) * H * W;
    const int groups = c_pad / 8;
    const size_t total = npix * groups;
    const size_t plane = static_cast<size_t>(H) * W;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int gch = idx % groups;
        const size_t pix = idx / groups;
        const size_t n = pix / plane, rem = pix % plane;
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = gch * 8 + 2 * j + e;
                v[e] = c < C ? x[(n * C + c) * plane + rem] : 0.f;
            }
            if (fmt == 1) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
            } else {
                __half2 h = __floats2half2_rn(v[0], v[1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
        }
        reinterpret_cast<uint4*>(out + pix * c_pad)[gch] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// NHWC u8 image -> NHWC 16-bit with channels zero-padded to c_pad: image_to_tensor (imgproc.py:1557: to_tensor = / 255 in
// fp32) fused with the layout change of the first convolution's input.
__global__ void u8nhwc_to_nhwc16_kernel(const unsigned char* __restrict__ x, uint16_t* __restrict__ out, size_t npix, int C, int c_pad,
                                        int fmt) {
    const int groups = c_pad / 8;
    const size_t total = npix * groups;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int gch = idx % groups;
        const size_t pix = idx / groups;
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = gch * 8 + 2 * j + e;
                v[e] = c < C ? __fdiv_rn(static_cast<float>(x[pix * C + c]), 255.f) : 0.f;
            }
            if (fmt == 1) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
            } else {
                __half2 h = __floats2half2_rn(v[0], v[1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
        }
        reinterpret_cast<uint4*>(out + pix * c_pad)[gch] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// imgproc.tensor_to_image (imgproc.py:1570-1596) on the device: NCHW float -> HWC u8 of image 0.
__global__ void tensor_to_image_kernel(const float* __restrict__ x, unsigned char* __restrict__ out, int C, size_t HW, int range_norm,
                                       int half) {
    const size_t total = HW * C;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = idx % C;
        const size_t p = idx / C;
        float v = x[static_cast<size_t>(c) * HW + p];
        if (range_norm) v = __fdiv_rn(__fadd_rn(v, 1.f), 2.f);            // imgproc.py:1587-1588
        if (half) {                                                        // :1591-1592: the arithmetic then runs in fp16
            const __half h = __hmul(__float2half_rn(v), __float2half_rn(255.f));
            v = fminf(fmaxf(__half2float(h), 0.f), 255.f);
        } else {
            v = fminf(fmaxf(__fmul_rn(v, 255.f), 0.f), 255.f);              // :1594
        }
        out[idx] = static_cast<unsigned char>(v);                           // astype("uint8"): truncation
    }
}

int grid_for(size_t total, int block) {
    size_t g = (total + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

// Packs every layer from the flat parameter vector: forward packs (+ padded biases) or the transposed data-gradient packs.
int launch_pack_all(resr_generator* g, const float* flat, int transposed, cudaStream_t s) {
    const Table& T = table();
    PackJob** slot = reinterpret_cast<PackJob**>(transposed ? &g->pack_jobs_t : &g->pack_jobs);
    if (!*slot) {
        std::vector<PackJob> h(kNumConvs);
        for (int k = 0; k < kNumConvs; ++k) {
            const ConvSpec& c = T.c[k];
            PackJob& j = h[k];
            j.p_off = c.p_off;
            j.cin = c.cin; j.cout = c.cout;
            if (!transposed) {
                j.w_off = c.w_off; j.b_off = c.b_off; j.nout = c.nout; j.nslices = c.nslices; j.nchunks = c.nchunks;
                j.fmt = g->precision == 1 ? 1 : c.fmt;
            } else {
                j.w_off = c.wt_off; j.b_off = 0; j.nout = 32; j.nslices = c.t_nslices; j.nchunks = c.t_nchunks; j.fmt = 1;
            }
        }
        if (cudaMalloc(slot, kNumConvs * sizeof(PackJob)) != cudaSuccess) return -1;
        cudaMemcpyAsync(*slot, h.data(), kNumConvs * sizeof(PackJob), cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);  // h is a local
    }
    // conv 0 (3 -> 64) never needs its input gradient: the transposed table is launched from layer 1
    const int first = transposed ? 1 : 0;
    pack_all_kernel<<<dim3(24, kNumConvs - first), 256, 0, s>>>(flat, *slot + first, transposed ? g->wpack_t : g->wpack, g->bias,
                                                                transposed);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// The mirrored dense-block data-gradient packs of all 69 blocks (needs the forward pack-job table on the device).
int launch_pack_rdb_bwd(resr_generator* g, const float* flat, uint8_t* wpack_t2, cudaStream_t s) {
    if (!g->pack_jobs) return -1;  // built by the forward pack (load_params always runs first)
    pack_rdb_bwd_kernel<<<dim3(8, 5, kNumRRDB * 3), 256, 0, s>>>(flat, static_cast<const PackJob*>(g->pack_jobs), wpack_t2);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

void launch_pack_conv(const float* w, const float* bias, uint16_t* wp, float* bp, int cin, int cout, int nout, int nslices,
                      int nchunks, int fmt, int transposed, cudaStream_t s) {
    const size_t total = static_cast<size_t>(nslices) * nchunks * 3 * (3 * nout) * 64;
    pack_conv_kernel<<<grid_for(total, 256), 256, 0, s>>>(w, bias, wp, bp, cin, cout, nout, nslices, nchunks, fmt, transposed);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
    size_t xin, c[3], f0, m[4], t1, t2, t3, t4, total;
};
static WsLayout ws_layout(size_t N, size_t H, size_t W, int precision = 0) {
    const size_t P = N * H * W;
    WsLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o = align_up(o + bytes, 1024);
        return at;
    };
    L.xin = take(P * 64 * 2);
    for (int i = 0; i < 3; ++i) L.c[i] = take(P * 192 * 2);
    L.f0 = take(P * 64 * 4);
    for (int i = 0; i < 4; ++i) L.m[i] = precision == 1 ? take(P * 64 * 4) : 0;  // bf16 recipe: fp32 masters of the residual stream
    L.t1 = take(4 * P * 64 * 2);
    L.t2 = take(16 * P * 64 * 2);
    L.t3 = take(16 * P * 64 * 2);
    L.t4 = take(16 * P * 64 * 2);
    L.total = o;
    return L;
}

static int build_plan(resr_generator* g, int N, int H, int W, void* ws, cudaStream_t s) {
    Plan& p = g->plan;
    p.valid = false;
    p.steps.clear();
    p.N = N; p.H = H; p.W = W; p.ws = ws;
    p.pair_policy = conv3x3_set_pair_policy(-1);
    const int prec = g->precision;
    const WsLayout L = ws_layout(N, H, W, prec);
    uint8_t* base = static_cast<uint8_t*>(ws);
    uint16_t* xin = reinterpret_cast<uint16_t*>(base + L.xin);
    // three 192-channel concat buffers: [0] holds the RRDB input (and receives the RRDB output in place), [1] / [2] the
    // inputs of its second / third dense block; F0 = fp32 copy of conv1's output for the global skip (model.py:261-262)
    uint16_t* cbuf[3];
    for (int i = 0; i < 3; ++i) cbuf[i] = reinterpret_cast<uint16_t*>(base + L.c[i]);
    float* F0 = reinterpret_cast<float*>(base + L.f0);
    float* M[4];
    for (int i = 0; i < 4; ++i) M[i] = reinterpret_cast<float*>(base + L.m[i]);
    uint16_t* t1 = reinterpret_cast<uint16_t*>(base + L.t1);
    uint16_t* t2 = reinterpret_cast<uint16_t*>(base + L.t2);
    uint16_t* t3 = reinterpret_cast<uint16_t*>(base + L.t3);
    uint16_t* t4 = reinterpret_cast<uint16_t*>(base + L.t4);
    p.xin = xin;

    struct Geo { int H, W, BW, BN, mode; };
    Geo geo[3];
    for (int s = 0; s < 3; ++s) {
        geo[s].H = H << s;
        geo[s].W = W << s;
        conv3x3_pick_tile(geo[s].W, &geo[s].BW, &geo[s].BN);
        geo[s].mode = geo[s].BN == 1 ? 0 : 1;
        if (g->force_mode >= 0) geo[s].mode = (geo[s].BN == 1) ? g->force_mode : 1;
    }
    // (one-time hygiene: no layer reads a channel that has not been written, the activation tensor maps end at Cin).
    // Issued on the CALLER's stream: ordered after whatever forward is still in flight on this workspace and capturable
    // into a CUDA graph (the legacy default stream is neither).
    for (int i = 0; i < 3; ++i) cudaMemsetAsync(cbuf[i], 0, static_cast<size_t>(N) * H * W * 192 * 2, s);

    const Table& T = table();
    int map_rc = 0;
    // in16: source activations [N, H<<s, W<<s, cin_total]
    auto make_step = [&](int conv, int s, const void* in16, int cin_total) {
        const ConvSpec& cs = T.c[conv];
        const Geo& q = geo[s];
        Step st;
        memset(&st, 0, sizeof(st));
        st.conv = conv;
        ConvArgs& a = st.a;
        a.N = N; a.H = q.H; a.W = q.W; a.BW = q.BW; a.BN = q.BN;
        a.nxs = (q.W + q.BW - 1) / q.BW;
        a.ncg = ((N + q.BN - 1) / q.BN) * a.nxs;
        a.nchunks = cs.nchunks;
        a.tail_ksteps = conv3x3_tail_ksteps(cs.cin);
        a.mode = q.mode;
        a.fmt_in = prec == 1 ? 1 : cs.fmt;
        a.rows_total = static_cast<long long>(a.ncg) * q.H;
        a.wpack = g->wpack + cs.w_off;
        a.bias = g->bias + cs.b_off;
        map_rc |= conv3x3_make_tmap_act(&st.maps.a, in16, N, q.H, q.W, cin_total, q.mode, q.BW, q.BN, cs.cin);
        return st;
    };
    auto set_out16 = [&](Step& st, int s, void* dst, int c_total, int choff, int fmt, int up2) {
        const Geo& q = geo[s];
        st.a.has_out16 = 1; st.a.out16_fmt = prec == 1 ? 1 : fmt; st.a.out16_choff = choff; st.a.out16_up2 = up2;
        map_rc |= conv3x3_make_tmap_out16(&st.maps.o16, dst, N, q.H, q.W, c_total, T.c[st.conv].nout, q.BW, q.BN, up2);
    };
    auto set_outf = [&](Step& st, float* dst) {
        st.a.has_outf = 1; st.a.outf_choff = 0; st.a.outf = static_cast<float*>(dst); st.a.outf_cstride = 64;
        map_rc |= conv3x3_make_tmap_f32(&st.maps.of, dst, N, geo[0].H, geo[0].W, 64, geo[0].BW, geo[0].BN);
    };
    auto set_res1 = [&](Step& st, const float* src) {
        st.a.has_res1 = 1; st.a.res_choff = 0; st.a.res1 = src; st.a.res1_cstride = 64;
    };
    auto push = [&](Step& st) {
        if (!conv3x3_choose(&st.a, T.c[st.conv].nout, T.c[st.conv].nslices, &st.cfg)) map_rc |= 1 << 20;
        // serpentine order: every other pair-kernel launch walks its rows / column groups back to front, so that it starts
        // on what the previous layer touched last (L2-resident) instead of on what it touched first (long evicted)
        static const int serp = getenv("RESR_CONV_SERPENTINE") ? atoi(getenv("RESR_CONV_SERPENTINE")) : 1;
        st.a.reverse = (serp && st.cfg.pair) ? static_cast<int>(p.steps.size() & 1) : 0;
        p.steps.push_back(st);
    };

    // 16-bit residual: the trunk's residual stream is the fp16 conv input itself (channels [0, 64) of a concat buffer)
    auto set_res16 = [&](Step& st, const uint16_t* src) {
        st.a.has_res1 = 1; st.a.res_choff = 0; st.a.res1 = src; st.a.res1_cstride = 192;
        st.a.res16 = 1; st.a.res16_fmt = 0;
    };
    int conv = 0;
    {   // conv1: model.py:258
        Step st = make_step(conv, 0, xin, 64);
        st.a.ep_mode = EP_PLAIN;
        set_outf(st, F0);
        set_out16(st, 0, cbuf[0], 192, 0, 0, 0);
        push(st);
        ++conv;
    }
    for (int i = 0; i < kNumRRDB; ++i) {
        for (int j = 0; j < 3; ++j) {
            uint16_t* cin_buf = cbuf[j];
            for (int k = 0; k < 4; ++k) {  // model.py:90-93
                Step st = make_step(conv, 0, cin_buf, 192);
                st.a.ep_mode = EP_PLAIN; st.a.lrelu = 1;
                set_out16(st, 0, cin_buf, 192, 64 + 32 * k, 0, 0);
                push(st);
                ++conv;
            }
            Step st = make_step(conv, 0, cin_buf, 192);  // model.py:94-96 (+ :129-130 for the third RDB)
            st.a.ep_mode = j < 2 ? EP_RDB : EP_RRDB;
            if (prec == 1) {
                // bf16 recipe (north_star): the residual stream is kept in fp32 masters next to the bf16 operands. Dense
                // block t reads master t % 4 (conv1's fp32 output for t = 0) and writes master (t + 1) % 4; the RRDB skip
                // reads the master of the RRDB input, which no block of the same RRDB overwrites.
                const int t = i * 3 + j;
                set_res1(st, t == 0 ? F0 : M[t % 4]);
                if (j == 2) { st.a.res2 = i == 0 ? F0 : M[(3 * i) % 4]; st.a.res2_cstride = 64; }
                set_outf(st, M[(t + 1) % 4]);
            } else {
                set_res16(st, cin_buf);
                if (j == 2) { st.a.res2 = cbuf[0]; st.a.res2_cstride = 192; }
            }
            set_out16(st, 0, cbuf[(j + 1) % 3], 192, 0, 0, 0);
            push(st);
            ++conv;
        }
    }
    {   // conv2 + skip (model.py:260-262), written nearest-upsampled x2 (model.py:264) in fp16
        Step st = make_step(conv, 0, cbuf[0], 192);
        st.a.ep_mode = EP_SKIP;
        set_res1(st, F0);
        set_out16(st, 0, t1, 64, 0, 0, 1);
        push(st);
        ++conv;
    }
    {   // upsampling1 conv + LeakyReLU (model.py:264), output written upsampled x2 again (model.py:265)
        Step st = make_step(conv, 1, t1, 64);
        st.a.lrelu = 1;
        set_out16(st, 1, t2, 64, 0, 0, 1);
        push(st);
        ++conv;
    }
    {   // upsampling2 conv + LeakyReLU (model.py:265)
        Step st = make_step(conv, 2, t2, 64);
        st.a.lrelu = 1;
        set_out16(st, 2, t3, 64, 0, 0, 0);
        push(st);
        ++conv;
    }
    {   // conv3 + LeakyReLU (model.py:267)
        Step st = make_step(conv, 2, t3, 64);
        st.a.lrelu = 1;
        set_out16(st, 2, t4, 64, 0, 0, 0);
        push(st);
        ++conv;
    }
    {   // conv4 + clamp (model.py:268-270): fp32 NCHW result, pointer patched per call
        Step st = make_step(conv, 2, t4, 64);
        st.a.clamp01 = 1;
        st.a.out_nchw = nullptr; st.a.out_nchw_c = 3;
        push(st);
        ++conv;
    }
    if (map_rc != 0) return set_error(RESR_E_CUDA, "tensor map / shared memory planning failed (%d)", map_rc);
    if (cudaGetLastError() != cudaSuccess) return set_error(RESR_E_CUDA, "workspace init failed");
    p.valid = true;
    return RESR_OK;
}

}  // namespace resr

using namespace resr;

extern "C" {

int resr_version(void) { return 1; }

size_t resr_generator_num_params(void) { return table().n_params; }
int resr_generator_num_tensors(void) { return 2 * kNumConvs; }
int resr_generator_launches_per_forward(void) { return 1 + kNumConvs; }

int resr_set_conv_pair_policy(int policy) { return conv3x3_set_pair_policy(policy); }

int resr_debug_wait_profile(unsigned long long* out16_host, int reset) {
    if (conv3x3_wait_profile(out16_host, reset) != 0) return set_error(RESR_E_CUDA, "wait profile copy failed");
    return RESR_OK;
}

int resr_generator_tensor_span(int index, size_t* offset, size_t* count) {
    if (index < 0 || index >= 2 * kNumConvs || !offset || !count) return set_error(RESR_E_INVALID, "bad tensor index");
    const ConvSpec& c = table().c[index / 2];
    const size_t wn = static_cast<size_t>(c.cout) * c.cin * 9;
    if (index % 2 == 0) { *offset = c.p_off; *count = wn; }
    else { *offset = c.p_off + wn; *count = c.cout; }
    return RESR_OK;
}

int resr_generator_create(resr_generator_t** out, int in_channels, int out_channels, int upscale_factor) {
    if (!out) return set_error(RESR_E_INVALID, "null out");
    if (in_channels != 3 || out_channels != 3 || upscale_factor != 4)
        return set_error(RESR_E_INVALID, "only Generator(3, 3, 4) is implemented (got %d, %d, %d)", in_channels,
                         out_channels, upscale_factor);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(RESR_E_CUDA, "no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return set_error(RESR_E_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return set_error(RESR_E_CUDA, "libresr needs an sm_100a device (found sm_%d%d)", prop.major, prop.minor);
    resr_generator* g = new resr_generator();
    g->num_sms = prop.multiProcessorCount;
    if (getenv("RESR_NUM_SMS") && atoi(getenv("RESR_NUM_SMS")) > 0 && atoi(getenv("RESR_NUM_SMS")) < g->num_sms)
        g->num_sms = atoi(getenv("RESR_NUM_SMS"));   // experiment knob: grids sized for a subset of the SMs (concurrent chains)
    const char* fm = getenv("RESR_CONV_MODE");
    if (fm) g->force_mode = atoi(fm);
    if (cudaMalloc(&g->wpack, table().pack_bytes) != cudaSuccess || cudaMalloc(&g->bias, table().bias_floats * 4) != cudaSuccess) {
        delete g;
        return set_error(RESR_E_CUDA, "cudaMalloc of packed weights failed");
    }
    *out = g;
    return RESR_OK;
}

void resr_generator_destroy(resr_generator_t* g) {
    if (!g) return;
    cudaFree(g->wpack);
    cudaFree(g->bias);
    cudaFree(g->wpack_t);
    cudaFree(g->wpack_t2);
    cudaFree(g->pack_jobs);
    cudaFree(g->pack_jobs_t);
    cudaFree(g->zero_bias);
    if (g->step_exec) cudaGraphExecDestroy(g->step_exec);
    if (g->h2d_stream) cudaStreamDestroy(g->h2d_stream);
    if (g->d2h_stream) cudaStreamDestroy(g->d2h_stream);
    for (int i = 0; i < 2; ++i) {
        if (g->ev_h2d[i]) cudaEventDestroy(g->ev_h2d[i]);
        if (g->ev_fwd[i]) cudaEventDestroy(g->ev_fwd[i]);
        if (g->ev_d2h[i]) cudaEventDestroy(g->ev_d2h[i]);
    }
    if (g->side_stream) cudaStreamDestroy(g->side_stream);
    if (g->ev_fork) cudaEventDestroy(g->ev_fork);
    if (g->ev_dy) cudaEventDestroy(g->ev_dy);
    if (g->ev_join) cudaEventDestroy(g->ev_join);
    for (int i = 0; i < 3; ++i)
        if (g->ev_dyc[i]) cudaEventDestroy(g->ev_dyc[i]);
    for (int i = 0; i < 4; ++i)
        if (g->ev_bucket[i]) cudaEventDestroy(g->ev_bucket[i]);
    delete g;
}

int resr_generator_load_params(resr_generator_t* g, const float* flat, void* stream) {
    if (!g || !flat) return set_error(RESR_E_INVALID, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Table& T = table();
    if (launch_pack_all(g, flat, 0, s) != 0) return set_error(RESR_E_CUDA, "weight packing failed");
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "pack kernels: %s", cudaGetErrorString(e));
    g->loaded = true;
    g->packed_t = false;
    g->flat_params = flat;
    ensure_transposed_packs(g, s, false);  // training handles keep the data-gradient packs in step with the weights
    return RESR_OK;
}

size_t resr_generator_workspace_bytes(int n, int h, int w) {
    if (n <= 0 || h <= 0 || w <= 0) return 0;
    return ws_layout(n, h, w).total;
}

size_t resr_generator_workspace_bytes_for(const resr_generator_t* g, int n, int h, int w) {
    if (!g || n <= 0 || h <= 0 || w <= 0) return 0;
    return ws_layout(n, h, w, g->precision).total;
}

int resr_generator_set_precision(resr_generator_t* g, int precision) {
    if (!g) return set_error(RESR_E_INVALID, "null argument");
    if (precision != 0 && precision != 1) return set_error(RESR_E_INVALID, "precision must be 0 (fp16) or 1 (bf16 + fp32 residual masters)");
    if (g->precision != precision) {
        g->precision = precision;
        g->loaded = false;        // the packed weights are in the old operand format: the caller reloads its parameters
        g->plan.valid = false;
        cudaFree(g->pack_jobs);   // the batched pack table carries the operand format
        g->pack_jobs = nullptr;
    }
    return RESR_OK;
}

static int forward_any(resr_generator_t* g, const float* x, const unsigned char* x_u8, float* y, unsigned char* y_u8, int n, int h, int w,
                       void* workspace, size_t workspace_bytes, void* stream);

int resr_generator_forward(resr_generator_t* g, const float* x, float* y, int n, int h, int w, void* workspace,
                           size_t workspace_bytes, void* stream) {
    if (!x || !y) return set_error(RESR_E_INVALID, "null argument");
    return forward_any(g, x, nullptr, y, nullptr, n, h, w, workspace, workspace_bytes, stream);
}

int resr_generator_forward_u8(resr_generator_t* g, const unsigned char* x_u8, unsigned char* y_u8, int n, int h, int w, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!x_u8 || !y_u8) return set_error(RESR_E_INVALID, "null argument");
    return forward_any(g, nullptr, x_u8, nullptr, y_u8, n, h, w, workspace, workspace_bytes, stream);
}

int resr_generator_forward_u8_host(resr_generator_t* g, const unsigned char* x_host, unsigned char* y_host, int n, int h, int w,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    if (!g || !x_host || !y_host) return set_error(RESR_E_INVALID, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t in_bytes = static_cast<size_t>(n) * 3 * h * w, out_bytes = in_bytes * 16;
    const size_t need = resr_generator_workspace_bytes_for(g, n, h, w);
    const size_t off_x = (need + 1023) / 1024 * 1024, off_y = off_x + (in_bytes + 1023) / 1024 * 1024;
    if (workspace_bytes < off_y + out_bytes) return set_error(RESR_E_NOMEM, "workspace too small for u8 host staging (need %zu)", off_y + out_bytes);
    unsigned char* xd = static_cast<unsigned char*>(workspace) + off_x;
    unsigned char* yd = static_cast<unsigned char*>(workspace) + off_y;
    if (cudaMemcpyAsync(xd, x_host, in_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) return set_error(RESR_E_CUDA, "H2D failed");
    const int rc = resr_generator_forward_u8(g, xd, yd, n, h, w, workspace, need, stream);
    if (rc != RESR_OK) return rc;
    if (cudaMemcpyAsync(y_host, yd, out_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess) return set_error(RESR_E_CUDA, "D2H failed");
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "forward_u8_host: %s", cudaGetErrorString(e));
    return RESR_OK;
}

int resr_tensor_to_image_u8(const float* x, unsigned char* out_hwc, int c, int h, int w, int range_norm, int half, void* stream) {
    if (!x || !out_hwc || c <= 0 || h <= 0 || w <= 0) return set_error(RESR_E_INVALID, "bad argument");
    const size_t HW = static_cast<size_t>(h) * w;
    tensor_to_image_kernel<<<grid_for(HW * c, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out_hwc, c, HW, range_norm, half);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "tensor_to_image: %s", cudaGetErrorString(e));
    return RESR_OK;
}

static int forward_any(resr_generator_t* g, const float* x, const unsigned char* x_u8, float* y, unsigned char* y_u8, int n, int h, int w,
                       void* workspace, size_t workspace_bytes, void* stream) {
    if (!g || !workspace) return set_error(RESR_E_INVALID, "null argument");
    if (n <= 0 || h <= 0 || w <= 0) return set_error(RESR_E_INVALID, "bad shape %dx3x%dx%d", n, h, w);
    if (!g->loaded) return set_error(RESR_E_INVALID, "resr_generator_load_params has not been called");
    if (workspace_bytes < resr_generator_workspace_bytes_for(g, n, h, w)) return set_error(RESR_E_NOMEM, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return set_error(RESR_E_INVALID, "workspace must be 1024-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Plan& p = g->plan;
    if (!p.valid || p.N != n || p.H != h || p.W != w || p.ws != workspace || p.pair_policy != conv3x3_set_pair_policy(-1)) {
        const int rc = build_plan(g, n, h, w, workspace, s);
        if (rc != RESR_OK) return rc;
    }
    const size_t total = static_cast<size_t>(n) * h * w * 8;
    if (x_u8) u8nhwc_to_nhwc16_kernel<<<grid_for(total, 256), 256, 0, s>>>(x_u8, p.xin, static_cast<size_t>(n) * h * w, 3, 64, g->precision == 1 ? 1 : 0);
    else nchw_to_nhwc16_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, p.xin, n, 3, h, w, 64, g->precision == 1 ? 1 : 0);
    const Table& T = table();
    for (size_t i = 0; i < p.steps.size(); ++i) {
        Step& st = p.steps[i];
        if (i + 1 == p.steps.size()) { st.a.out_nchw = y; st.a.out_u8 = y_u8; }
        const cudaError_t e = conv3x3_run(st.maps, st.a, st.cfg, g->num_sms, s);
        if (e != cudaSuccess) return set_error(RESR_E_CUDA, "conv %d launch: %s", st.conv, cudaGetErrorString(e));
    }
    return RESR_OK;
}

int resr_generator_forward_host(resr_generator_t* g, const float* x_host, float* y_host, int n, int h, int w,
                                void* workspace, size_t workspace_bytes, void* stream) {
    if (!x_host || !y_host) return set_error(RESR_E_INVALID, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t in_bytes = static_cast<size_t>(n) * 3 * h * w * 4, out_bytes = in_bytes * 16;
    const size_t need = resr_generator_workspace_bytes_for(g, n, h, w);
    // device staging for x and y lives at the end of the caller's workspace
    const size_t off_x = (need + 1023) / 1024 * 1024, off_y = off_x + (in_bytes + 1023) / 1024 * 1024;
    if (workspace_bytes < off_y + out_bytes) return set_error(RESR_E_NOMEM, "workspace too small for host staging (need %zu)", off_y + out_bytes);
    float* xd = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + off_x);
    float* yd = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + off_y);
    if (cudaMemcpyAsync(xd, x_host, in_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) return set_error(RESR_E_CUDA, "H2D failed");
    const int rc = resr_generator_forward(g, xd, yd, n, h, w, workspace, need, stream);
    if (rc != RESR_OK) return rc;
    if (cudaMemcpyAsync(y_host, yd, out_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess) return set_error(RESR_E_CUDA, "D2H failed");
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "forward_host: %s", cudaGetErrorString(e));
    return RESR_OK;
}

int resr_generator_forward_host_async(resr_generator_t* g, const float* x_host, float* y_host, int n, int h, int w,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    if (!g || !x_host || !y_host) return set_error(RESR_E_INVALID, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t in_bytes = static_cast<size_t>(n) * 3 * h * w * 4, out_bytes = in_bytes * 16;
    const size_t need = resr_generator_workspace_bytes_for(g, n, h, w);
    const size_t in_al = (in_bytes + 1023) / 1024 * 1024, out_al = (out_bytes + 1023) / 1024 * 1024;
    const size_t off0 = (need + 1023) / 1024 * 1024;
    if (workspace_bytes < off0 + 2 * (in_al + out_al))
        return set_error(RESR_E_NOMEM, "workspace too small for two host staging slots (need %zu)", off0 + 2 * (in_al + out_al));
    if (!g->h2d_stream) {
        if (cudaStreamCreateWithFlags(&g->h2d_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&g->d2h_stream, cudaStreamNonBlocking) != cudaSuccess)
            return set_error(RESR_E_CUDA, "cannot create copy streams");
        for (int i = 0; i < 2; ++i) {
            cudaEventCreateWithFlags(&g->ev_h2d[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&g->ev_fwd[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&g->ev_d2h[i], cudaEventDisableTiming);
        }
    }
    if (g->host_calls > 0 && (g->host_shape[0] != n || g->host_shape[1] != h || g->host_shape[2] != w || g->host_ws != workspace)) {
        // a different shape (or workspace) moves the staging slots: drain every queued copy / forward before they are reused
        cudaStreamSynchronize(g->h2d_stream);
        cudaStreamSynchronize(s);
        cudaStreamSynchronize(g->d2h_stream);
        g->host_calls = 0;
    }
    g->host_shape[0] = n; g->host_shape[1] = h; g->host_shape[2] = w; g->host_ws = workspace;
    const int slot = static_cast<int>(g->host_calls & 1);
    const bool reused = g->host_calls >= 2;  // this slot has been through a full cycle before
    uint8_t* base = static_cast<uint8_t*>(workspace) + off0 + static_cast<size_t>(slot) * (in_al + out_al);
    float* xd = reinterpret_cast<float*>(base);
    float* yd = reinterpret_cast<float*>(base + in_al);
    // H2D into the slot once the forward that last read it is done
    if (reused) cudaStreamWaitEvent(g->h2d_stream, g->ev_fwd[slot], 0);
    if (cudaMemcpyAsync(xd, x_host, in_bytes, cudaMemcpyHostToDevice, g->h2d_stream) != cudaSuccess) return set_error(RESR_E_CUDA, "H2D failed");
    cudaEventRecord(g->ev_h2d[slot], g->h2d_stream);
    // forward once the input has landed and the slot's previous result has left the device
    cudaStreamWaitEvent(s, g->ev_h2d[slot], 0);
    if (reused) cudaStreamWaitEvent(s, g->ev_d2h[slot], 0);
    const int rc = resr_generator_forward(g, xd, yd, n, h, w, workspace, need, stream);
    if (rc != RESR_OK) return rc;
    cudaEventRecord(g->ev_fwd[slot], s);
    cudaStreamWaitEvent(g->d2h_stream, g->ev_fwd[slot], 0);
    if (cudaMemcpyAsync(y_host, yd, out_bytes, cudaMemcpyDeviceToHost, g->d2h_stream) != cudaSuccess) return set_error(RESR_E_CUDA, "D2H failed");
    cudaEventRecord(g->ev_d2h[slot], g->d2h_stream);
    ++g->host_calls;
    return RESR_OK;
}

int resr_generator_host_sync(resr_generator_t* g) {
    if (!g) return set_error(RESR_E_INVALID, "null argument");
    if (!g->d2h_stream) return RESR_OK;
    cudaError_t e = cudaStreamSynchronize(g->h2d_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->d2h_stream);
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "host_sync: %s", cudaGetErrorString(e));
    return RESR_OK;
}

int resr_nchw_to_nhwc16(const float* x, void* out16, int n, int c, int h, int w, int c_pad, int fmt, void* stream) {
    if (!x || !out16 || c_pad % 8 != 0 || c > c_pad) return set_error(RESR_E_INVALID, "bad argument");
    const size_t total = static_cast<size_t>(n) * h * w * (c_pad / 8);
    nchw_to_nhwc16_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, static_cast<uint16_t*>(out16), n, c, h, w, c_pad, fmt);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "nchw_to_nhwc16: %s", cudaGetErrorString(e));
    return RESR_OK;
}

int resr_conv3x3(const resr_conv_desc* d, void* stream) {
    if (!d || !d->in16 || !d->weight) return set_error(RESR_E_INVALID, "null argument");
    if (d->cin > d->c_total || d->c_total % 8 != 0) return set_error(RESR_E_INVALID, "bad channel counts");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nout = d->cout >= 32 ? 32 : 16;
    const int nslices = (d->cout + nout - 1) / nout;
    const int nchunks = (d->cin + 63) / 64;
    if (nchunks * 64 > d->c_total) return set_error(RESR_E_INVALID, "c_total must cover cin rounded up to 64");
    const size_t pack_bytes = static_cast<size_t>(nslices) * nchunks * 3 * (3 * nout) * 128;
    uint8_t* wp = nullptr;
    float* bp = nullptr;
    if (cudaMalloc(&wp, pack_bytes) != cudaSuccess || cudaMalloc(&bp, nslices * nout * 4) != cudaSuccess)
        return set_error(RESR_E_CUDA, "cudaMalloc failed");
    launch_pack_conv(d->weight, d->bias, reinterpret_cast<uint16_t*>(wp), bp, d->cin, d->cout, nout, nslices, nchunks,
                     d->fmt_in, 0, s);
    cudaStreamSynchronize(s);  // the conv prefetches its weights before the programmatic grid dependency resolves
    ConvArgs a;
    ConvMaps maps;
    memset(&a, 0, sizeof(a));
    memset(&maps, 0, sizeof(maps));
    a.N = d->n; a.H = d->h; a.W = d->w;
    conv3x3_pick_tile(d->w, &a.BW, &a.BN);
    a.mode = d->mode >= 0 ? d->mode : (a.BN == 1 ? 0 : 1);
    if (a.BN != 1) a.mode = 1;
    a.nxs = (d->w + a.BW - 1) / a.BW;
    a.ncg = ((d->n + a.BN - 1) / a.BN) * a.nxs;
    a.nchunks = nchunks;
    a.tail_ksteps = conv3x3_tail_ksteps(d->cin);
    a.fmt_in = d->fmt_in;
    a.rows_total = static_cast<long long>(a.ncg) * d->h;
    a.wpack = wp; a.bias = bp;
    a.ep_mode = d->ep_mode; a.lrelu = d->lrelu; a.clamp01 = d->clamp01;
    int rc = conv3x3_make_tmap_act(&maps.a, d->in16, d->n, d->h, d->w, d->c_total, a.mode, a.BW, a.BN, d->cin);
    if (d->out16) {
        a.has_out16 = 1; a.out16_fmt = d->out16_fmt; a.out16_choff = d->out16_choff; a.out16_up2 = d->out16_up2;
        rc |= conv3x3_make_tmap_out16(&maps.o16, d->out16, d->n, d->h, d->w, d->out16_cstride, nout, a.BW, a.BN, d->out16_up2);
    }
    if (d->outf) {
        a.has_outf = 1; a.outf_choff = d->outf_choff; a.outf = static_cast<float*>(d->outf); a.outf_cstride = d->outf_cstride;
        rc |= conv3x3_make_tmap_f32(&maps.of, d->outf, d->n, d->h, d->w, d->outf_cstride, a.BW, a.BN);
    }
    if (d->res1) {
        a.has_res1 = 1; a.res_choff = d->res_choff; a.res1 = d->res1; a.res1_cstride = d->res_cstride;
        a.res16 = d->res16; a.res16_fmt = d->res16_fmt;
    }
    a.res2 = d->res2; a.res2_cstride = d->res_cstride;
    a.out_nchw = d->out_nchw; a.out_nchw_c = d->out_nchw_c;
    a.dbg = d->dbg;
    a.dbg_flags = d->dbg_flags;
    ConvLaunchCfg cfg;
    if (!conv3x3_choose(&a, nout, nslices, &cfg)) rc |= 1 << 20;
    if (nout == 16 && (d->out16 || d->outf || d->res1)) rc |= 1 << 21;  // the 16-wide slice only feeds the NCHW output
    cudaError_t e = cudaSuccess;
    if (rc == 0) e = conv3x3_run(maps, a, cfg, sms, s);
    const cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(wp);
    cudaFree(bp);
    if (rc != 0) return set_error(RESR_E_CUDA, "tensor map / shared memory planning failed (%d)", rc);
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "conv launch: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return set_error(RESR_E_CUDA, "conv run: %s", cudaGetErrorString(e2));
    return RESR_OK;
}

}  // extern "C"
