// C ABI of the MBExWN forward pass (include/mbexwn.h): handle, tensor registry, workspace carving and the stage
// orchestration of MBExWN.call (custom_pulsed_generator.py:556-771) on the padded frame grid.
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/mbexwn.h"
#include "kernels.cuh"
#include "wn_tc.cuh"

// One WaveNetAEBlock as the engine runs it: the model config with the block's own WaveNet geometry substituted
// (wn_c, wn_cin, wn_cond_conv_up, wn_name, steps_per_frame), so every WaveNet routine takes a block like a single-block model.
struct WnBlock {
    mbexwn_config_t cfg;
    int up = 1;                 // sub-pixel up-sampling conv after the block (custom_AE_layers.py:519-526), 1 = none
    std::string up_name;
};

struct mbexwn_handle_s {
    mbexwn_config_t cfg;
    std::vector<WnBlock> blocks;    // >= 1 entry
    int out_steps = 0;              // rows per frame behind the last block = hop / subbands
    int device = 0;
    std::map<std::string, std::pair<const void*, size_t>> tensors;
    std::string error;
    int launches = 0;
    int debug_taps = 1;
    int stage_timing = 0;
    int stop_after_f0 = 0;      // option: F0 pass of chunked long-form synthesis
    int tc_subnets = 1;         // wide sub-net / conditioning convs on the tensor cores (TC precisions only)
    int fuse_tail = 1;          // LinInterp -> 1x1 -> LinInterp tail of a sub-net as one launch
    std::map<std::string, float> host_scalars;   // one-element tensors whose value the host knows (mbexwn_set_scalar)
    cudaEvent_t ev[MBEXWN_N_STAGES + 1] = {};
    bool ev_ready = false;
    bool ev_recorded = false;
    mbx::WnTcState tc;
    // pipelined host-buffer forward (mbexwn_forward_host_begin / _wait): two copy streams and per-slot events
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_h2d[2] = {}, ev_done[2] = {}, ev_d2h[2] = {};
    bool pipe_ready = false;
    bool slot_busy[2] = {false, false};
    int* range_flag = nullptr;   // host-mapped word the f16f8 kernels OR their range findings into (mbexwn_range_status)
};

namespace mbx {

int set_error(mbexwn_handle_t h, const char* what, cudaError_t e) {
    if (h) {
        h->error = std::string(what) + ": " + cudaGetErrorString(e);
    }
    return MBEXWN_ERR_CUDA;
}

static int fail(mbexwn_handle_t h, int code, const std::string& msg) {
    if (h) h->error = msg;
    return code;
}

// ---- workspace ----------------------------------------------------------------------------------------
struct Slot {
    size_t off = 0, bytes = 0;
};

struct Workspace {
    std::map<std::string, Slot> slots;
    size_t total = 0;
    void add(const std::string& name, size_t bytes) {
        size_t aligned = (bytes + 1023) & ~size_t(1023);
        slots[name] = Slot{total, bytes};
        total += aligned;
    }
};

static size_t subnet_scratch_elems(const mbexwn_op_t* ops, int n, int n_mel) {
    size_t m = (size_t)n_mel;
    for (int i = 0; i < n; ++i) {
        size_t a = (size_t)ops[i].rate_out * ops[i].ch_out;
        if (a > m) m = a;
    }
    return m;
}

static int round64(int x) { return (x + 63) / 64 * 64; }

// A conv op runs on the tensor cores when it is wide enough to fill MMA tiles and its epilogue is one the tap-GEMM has
static bool tc_eligible(const mbexwn_op_t& op) {
    if (op.kind != 0 || op.cin < 32 || op.cout < 16 || op.cout % 8 || op.act > ACT_LEAKY || op.k > 16) return false;
    const int f = op.subpixel > 0 ? op.subpixel : 1;
    return op.cout % f == 0 && (op.cout / f) % 8 == 0 && (op.dilation <= 1);
}

static size_t subnet_hilo_elems(const mbexwn_op_t* ops, int n) {
    size_t m = 0;
    for (int i = 0; i < n; ++i) {
        if (!tc_eligible(ops[i])) continue;
        size_t a = (size_t)ops[i].rate_out * 2 * round64(ops[i].ch_out);
        if (a > m) m = a;
    }
    return m;
}

static Workspace carve(const mbexwn_handle_s& hd, int64_t F, int64_t n_chunks, int precision, int debug_taps) {
    const mbexwn_config_t& c = hd.cfg;
    Workspace w;
    const size_t f4 = sizeof(float);
    const int64_t rows = F * c.steps_per_frame;             // rate of the WaveNet input (block 0)
    const int64_t out_rows = F * hd.out_steps;              // rate of the sub-band signals
    size_t sn = subnet_scratch_elems(c.pp_ops, c.n_pp_ops, c.mel_channels);
    size_t sn2 = subnet_scratch_elems(c.ps_ops, c.n_ps_ops, c.mel_channels);
    if (sn2 > sn) sn = sn2;
    w.add("sn_a", (size_t)F * sn * f4);
    w.add("sn_b", (size_t)F * sn * f4);
    if (precision != MBEXWN_PREC_FP32_SIMT) {
        // bf16 [hi | lo] activations of the tensor-core sub-net convs: packed mel + ping-pong layer outputs
        size_t hl = subnet_hilo_elems(c.pp_ops, c.n_pp_ops), hl2 = subnet_hilo_elems(c.ps_ops, c.n_ps_ops);
        if (hl2 > hl) hl = hl2;
        w.add("mel_hl", (size_t)F * 2 * round64(c.mel_channels) * 2);
        w.add("sn_h0", (size_t)F * hl * 2);
        w.add("sn_h1", (size_t)F * hl * 2);
    }
    w.add("F0", (size_t)F * c.pulse_per_frame * f4);
    w.add("cum", (size_t)F * c.pulse_per_frame * f4);
    w.add("chunk_off", (size_t)(n_chunks + 1) * f4);
    if (debug_taps || c.pulse_pqmf_taps > 0) w.add("pulse", (size_t)F * c.pulse_per_frame * f4);   // input of the pulse PQMF analysis
    if (debug_taps) {
        w.add("phase", (size_t)F * c.pulse_per_frame * f4);
        w.add("index", (size_t)F * c.pulse_per_frame * sizeof(int32_t));
        w.add("vtf", (size_t)F * (c.fft_size / 2 + 1) * 2 * f4);
        w.add("lifter_index", (size_t)F * sizeof(int32_t));
    }
    w.add("wn_in", (size_t)rows * c.wn_cin * f4);
    {
        // the blocks run one after the other and share their buffers: every slot takes its largest user
        std::vector<std::pair<std::string, size_t>> wn;
        auto need = [&](const std::string& name, size_t bytes) {
            for (auto& e : wn)
                if (e.first == name) { if (bytes > e.second) e.second = bytes; return; }
            wn.emplace_back(name, bytes);
        };
        const int out_pad = wn_tc_out_pad(c);
        for (size_t ib = 0; ib < hd.blocks.size(); ++ib) {
            const mbexwn_config_t& bc = hd.blocks[ib].cfg;
            const int64_t brows = F * bc.steps_per_frame;
            need("cond", (size_t)F * bc.wn_cond_conv_up * 2 * bc.wn_c * f4);
            need("wn_out", (size_t)brows * out_pad * f4);
            if (precision == MBEXWN_PREC_FP32_SIMT) {
                need("skip", (size_t)brows * bc.wn_c * f4);
                need("h", (size_t)brows * bc.wn_c * f4);
                need("z", (size_t)brows * 2 * bc.wn_c * f4);
                need("act", (size_t)brows * bc.wn_c * f4);
                need("rs", (size_t)brows * 2 * bc.wn_c * f4);
            } else {
                wn_tc_carve(bc, brows, precision, [&](const char* n, size_t b) { need(n, b); });
            }
            // up-sampling conv output: the next block's input (rows x wn_cout) or, behind the last block, the post net's
            // input with the row pitch of wn_out
            if (hd.blocks[ib].up > 1)
                need("blk_in", (size_t)brows * hd.blocks[ib].up * (ib + 1 < hd.blocks.size() ? c.wn_cout : out_pad) * f4);
        }
        for (auto& e : wn) w.add(e.first, e.second);
    }
    w.add("subbands", (size_t)out_rows * c.subbands * f4);
    w.add("excitation", (size_t)F * c.hop * f4);
    w.add("ceps", (size_t)F * c.n_ceps * f4);
    w.add("frames", (size_t)F * c.stft_win * f4);
    if (c.norm_enable) {
        w.add("mel_norm", (size_t)F * c.mel_channels * f4);
        w.add("norm_rms_a", (size_t)F * f4);
        w.add("norm_rms_b", (size_t)F * f4);
        if (debug_taps) w.add("norm_gain", (size_t)F * c.hop * f4);
    }
    return w;
}

struct Ctx {
    mbexwn_handle_t h;
    const mbexwn_batch_t* b;
    FrameGrid g;
    Workspace ws;
    char* base;
    cudaStream_t s;
    template <class T>
    T* p(const char* name) {
        auto it = ws.slots.find(name);
        return it == ws.slots.end() ? nullptr : reinterpret_cast<T*>(base + it->second.off);
    }
};

static const float* tensor(mbexwn_handle_t h, const std::string& name, size_t min_bytes, int* rc) {
    auto it = h->tensors.find(name);
    if (it == h->tensors.end()) {
        *rc = fail(h, MBEXWN_ERR_MISSING, "tensor not registered: " + name);
        return nullptr;
    }
    if (it->second.second < min_bytes) {
        *rc = fail(h, MBEXWN_ERR_INVALID, "tensor too small: " + name);
        return nullptr;
    }
    return reinterpret_cast<const float*>(it->second.first);
}

static ConvArgs conv_args(const mbexwn_config_t& c, const mbexwn_op_t& op, int rate, long long rows,
                          const float* x, const float* w, const float* bias, const float* alpha, float* out) {
    ConvArgs a{};
    a.x = x; a.ld_x = op.cin; a.w = w; a.bias = bias; a.alpha = alpha; a.out = out; a.ld_out = op.cout;
    a.rows = rows; a.rate = rate;
    a.k = op.k; a.cin = op.cin; a.cout = op.cout; a.dilation = op.dilation > 0 ? op.dilation : 1;
    a.pad_l = op.pad_l; a.pad_mode = op.pad_mode;
    a.act = op.act; a.act_mod = op.act_channels > 0 ? op.act_channels : op.cout;
    a.leaky = c.leaky_alpha; a.a0 = c.f0_span; a.a1 = c.f0_min;
    return a;
}

// ops[i .. i + 2] = LinInterp + act, 1x1 conv to one channel, LinInterp + act ending the program: one fused launch.
// Returns 1 if it ran, 0 if the pattern does not apply, < 0 on error.
static int try_subnet_tail(Ctx& cx, const mbexwn_op_t* ops, int n_ops, int i, const float* cur, float* final_out) {
    mbexwn_handle_t h = cx.h;
    const mbexwn_config_t& c = h->cfg;
    if (!h->fuse_tail || i + 2 != n_ops - 1) return 0;
    const mbexwn_op_t &l1 = ops[i], &cv = ops[i + 1], &l2 = ops[i + 2];
    if (l1.kind != 1 || cv.kind != 0 || l2.kind != 1 || cv.k != 1 || cv.cout != 1 || cv.act != ACT_NONE ||
        cv.cin != l1.ch_out || cv.subpixel > 1 || cv.rate_in != l1.rate_out || l2.rate_in != cv.rate_out)
        return 0;
    int rc = 0;
    SubnetTailArgs a{};
    a.x = cur; a.out = final_out; a.rows_in = (long long)cx.g.n_frames * l1.rate_in; a.rate_in = l1.rate_in;
    a.ch = l1.ch_out; a.up1 = l1.up; a.up2 = l2.up; a.act1 = l1.act; a.act2 = l2.act;
    a.leaky = c.leaky_alpha; a.a0 = c.f0_span; a.a1 = c.f0_min;
    if (!subnet_tail_supported(a)) return 0;
    if (l1.act == ACT_PRELU) {
        a.alpha1 = tensor(h, std::string(l1.act_name) + "/alpha", (size_t)l1.act_channels * 4, &rc);
        if (!a.alpha1) return rc;
    }
    a.w = tensor(h, std::string(cv.name) + "/W", (size_t)cv.cin * 4, &rc);
    if (!a.w) return rc;
    auto it = h->host_scalars.find(std::string(cv.name) + "/b");
    if (it == h->host_scalars.end()) return 0;               // bias value not known on the host: keep the three launches
    a.bias = it->second;
    MBX_CUDA_CHECK(launch_subnet_tail(a, cx.g, cx.s));
    h->launches++;
    return 1;
}

// Run one mel-rate sub-net program; the last op writes into `final_out`.
static int run_subnet(Ctx& cx, const mbexwn_op_t* ops, int n_ops, const float* input, float* final_out) {
    mbexwn_handle_t h = cx.h;
    const mbexwn_config_t& c = h->cfg;
    float* bufs[2] = {cx.p<float>("sn_a"), cx.p<float>("sn_b")};
    const float* cur = input;
    int flip = 0;
    for (int i = 0; i < n_ops; ++i) {
        const mbexwn_op_t& op = ops[i];
        float* out = (i == n_ops - 1) ? final_out : bufs[flip];
        int rc = 0;
        if (op.kind == 1) {
            rc = try_subnet_tail(cx, ops, n_ops, i, cur, final_out);
            if (rc < 0) return rc;
            if (rc == 1) return MBEXWN_OK;
            rc = 0;
        }
        const float* alpha = nullptr;
        if (op.act == ACT_PRELU) {
            alpha = tensor(h, std::string(op.act_name) + "/alpha", (size_t)op.act_channels * 4, &rc);
            if (!alpha) return rc;
        }
        if (op.kind == 0) {
            const float* w = tensor(h, std::string(op.name) + "/W", (size_t)op.k * op.cin * op.cout * 4, &rc);
            if (!w) return rc;
            const float* bias = tensor(h, std::string(op.name) + "/b", (size_t)op.cout * 4, &rc);
            if (!bias) return rc;
            ConvArgs a = conv_args(c, op, op.rate_in, (long long)cx.g.n_frames * op.rate_in, cur, w, bias, alpha, out);
            MBX_CUDA_CHECK(launch_conv1d(a, cx.g, cx.s));
        } else {
            LinInterpArgs a{};
            a.x = cur; a.out = out; a.rows_in = (long long)cx.g.n_frames * op.rate_in; a.rate_in = op.rate_in;
            a.ch = op.ch_out; a.up = op.up; a.act = op.act; a.alpha = alpha; a.leaky = c.leaky_alpha;
            a.a0 = c.f0_span; a.a1 = c.f0_min;
            MBX_CUDA_CHECK(launch_lininterp(a, cx.g, cx.s));
        }
        h->launches++;
        cur = out;
        flip ^= 1;
    }
    return MBEXWN_OK;
}

// Tensor-core variant of run_subnet: wide convs run as 3-product bf16 tap-GEMMs whose epilogue writes the bf16
// [hi | lo] planes the next conv reads (fp32 only where a CUDA-core op or the caller consumes the result).
static int run_subnet_tc(Ctx& cx, const mbexwn_op_t* ops, int n_ops, const float* input, float* final_out) {
    mbexwn_handle_t h = cx.h;
    const mbexwn_config_t& c = h->cfg;
    float* bufs[2] = {cx.p<float>("sn_a"), cx.p<float>("sn_b")};
    char* hbufs[2] = {cx.p<char>("sn_h0"), cx.p<char>("sn_h1")};
    const float* cur = input;        // fp32 view of the current activations (nullptr while they only exist as hi/lo)
    const char* cur_hl = nullptr;    // bf16 [hi | lo] view
    int cur_cpad = 0, flip = 0, hflip = 0;
    for (int i = 0; i < n_ops; ++i) {
        const mbexwn_op_t& op = ops[i];
        const bool last = i == n_ops - 1;
        int rc = 0;
        const float* alpha = nullptr;
        if (op.act == ACT_PRELU) {
            alpha = tensor(h, std::string(op.act_name) + "/alpha", (size_t)op.act_channels * 4, &rc);
            if (!alpha) return rc;
        }
        if (tc_eligible(op) && (cur_hl || i == 0)) {
            const int cin_pad = round64(op.cin);
            if (op.pad_mode != PAD_ZERO && c.halo_frames * op.rate_in < op.pad_l + op.pad_r)
                return fail(h, MBEXWN_ERR_INVALID, "halo_frames too small for the mirrored pad rows of the sub-net convs");
            if (!cur_hl) {                                   // first layer: pack the fp32 input with its pad rows
                MBX_RC(wn_tc_pack(cur, cx.p<char>("mel_hl"), (long long)cx.g.n_frames * op.rate_in, op.cin, cin_pad, op.rate_in,
                                  op.pad_l, op.pad_r, op.pad_mode, cx.g, cx.s, &h->error));
                cur_hl = cx.p<char>("mel_hl");
                cur_cpad = cin_pad;
                h->launches++;
            }
            if (cur_cpad != cin_pad) return fail(h, MBEXWN_ERR_INVALID, "sub-net channel mismatch");
            const float* w = tensor(h, std::string(op.name) + "/tc/W", (size_t)op.cout * 2 * op.k * cin_pad * 2, &rc);
            if (!w) return rc;
            const float* bias = tensor(h, std::string(op.name) + "/b", (size_t)op.cout * 4, &rc);
            if (!bias) return rc;
            const int f = op.subpixel > 0 ? op.subpixel : 1;
            // the next conv reads hi/lo planes only if they need no channel padding (cout / f a multiple of 64)
            const bool next_tc = !last && tc_eligible(ops[i + 1]) && (op.cout / f) % 64 == 0;
            TcConvArgs a{};
            a.a_hilo = cur_hl; a.rows = (long long)cx.g.n_frames * op.rate_in; a.cin_pad = cin_pad;
            a.w = w; a.cout = op.cout; a.k = op.k; a.dilation = 1; a.pad_l = op.pad_l;
            a.bias = bias; a.act = op.act; a.alpha = alpha; a.act_mod = op.act_channels > 0 ? op.act_channels : op.cout;
            a.leaky = c.leaky_alpha; a.rate = op.rate_in;
            if (next_tc) {
                a.out_hilo = hbufs[hflip]; a.out_cpad = round64(op.cout / f); a.subpixel = f;
            } else {
                a.out_f32 = last ? final_out : bufs[flip]; a.ld_out = op.cout;       // (T, cout) == (T f, cout / f)
            }
            MBX_RC(wn_tc_conv(h->tc, a, cx.g, cx.s, &h->error));
            h->launches++;
            if (next_tc) {
                const mbexwn_op_t& nx = ops[i + 1];
                if (nx.pad_mode != PAD_ZERO) {
                    MBX_RC(wn_tc_mirror(hbufs[hflip], (long long)cx.g.n_frames * op.rate_out, a.out_cpad, op.rate_out, nx.pad_l,
                                        nx.pad_r, nx.pad_mode, cx.g, cx.s, &h->error));
                    h->launches++;
                }
                cur_hl = hbufs[hflip]; cur_cpad = a.out_cpad; cur = nullptr;
                hflip ^= 1;
            } else {
                cur = last ? final_out : bufs[flip]; cur_hl = nullptr;
                flip ^= 1;
            }
            continue;
        }
        if (!cur) return fail(h, MBEXWN_ERR_INVALID, "sub-net op needs fp32 activations");
        if (op.kind == 1) {
            rc = try_subnet_tail(cx, ops, n_ops, i, cur, final_out);
            if (rc < 0) return rc;
            if (rc == 1) return MBEXWN_OK;
            rc = 0;
        }
        float* out = last ? final_out : bufs[flip];
        if (op.kind == 0) {
            const float* w = tensor(h, std::string(op.name) + "/W", (size_t)op.k * op.cin * op.cout * 4, &rc);
            if (!w) return rc;
            const float* bias = tensor(h, std::string(op.name) + "/b", (size_t)op.cout * 4, &rc);
            if (!bias) return rc;
            ConvArgs a = conv_args(c, op, op.rate_in, (long long)cx.g.n_frames * op.rate_in, cur, w, bias, alpha, out);
            MBX_CUDA_CHECK(launch_conv1d(a, cx.g, cx.s));
        } else {
            LinInterpArgs a{};
            a.x = cur; a.out = out; a.rows_in = (long long)cx.g.n_frames * op.rate_in; a.rate_in = op.rate_in;
            a.ch = op.ch_out; a.up = op.up; a.act = op.act; a.alpha = alpha; a.leaky = c.leaky_alpha;
            a.a0 = c.f0_span; a.a1 = c.f0_min;
            MBX_CUDA_CHECK(launch_lininterp(a, cx.g, cx.s));
        }
        h->launches++;
        cur = out;
        flip ^= 1;
    }
    return MBEXWN_OK;
}

static mbexwn_op_t simple_conv(int k, int cin, int cout, int dil, int pad_l) {
    mbexwn_op_t op{};
    op.kind = 0; op.k = k; op.cin = cin; op.cout = cout; op.dilation = dil; op.pad_l = pad_l; op.pad_r = pad_l;
    op.pad_mode = PAD_ZERO; op.act = ACT_NONE; op.act_channels = cout; op.subpixel = 1;
    return op;
}

static int wavenet_fp32(Ctx& cx, const mbexwn_config_t& c, const float* wn_in, int ld_in) {
    mbexwn_handle_t h = cx.h;
    const long long rows = (long long)cx.g.n_frames * c.steps_per_frame;
    const std::string n = c.wn_name;
    int rc = 0;
    float *hbuf = cx.p<float>("h"), *z = cx.p<float>("z"), *act = cx.p<float>("act"), *rs = cx.p<float>("rs");
    float* skip = cx.p<float>("skip");
    const float* cond = cx.p<float>("cond");
    {   // start 1x1 (custom_AE_layers.py:280)
        const float* w = tensor(h, n + "/start/W", (size_t)c.wn_cin * c.wn_c * 4, &rc); if (!w) return rc;
        const float* b = tensor(h, n + "/start/b", (size_t)c.wn_c * 4, &rc); if (!b) return rc;
        mbexwn_op_t op = simple_conv(1, c.wn_cin, c.wn_c, 1, 0);
        ConvArgs sa = conv_args(c, op, c.steps_per_frame, rows, wn_in, w, b, nullptr, hbuf);
        sa.ld_x = ld_in;
        MBX_CUDA_CHECK(launch_conv1d(sa, cx.g, cx.s));
        h->launches++;
    }
    for (int i = 0; i < c.wn_layers; ++i) {
        const int d = c.wn_dilations[i];
        const int n_rs = (i < c.wn_layers - 1) ? 2 * c.wn_c : c.wn_c;
        const std::string li = std::to_string(i);
        const float* w = tensor(h, n + "/conv1D_" + li + "/W", (size_t)c.wn_k * c.wn_c * 2 * c.wn_c * 4, &rc); if (!w) return rc;
        const float* b = tensor(h, n + "/conv1D_" + li + "/b", (size_t)2 * c.wn_c * 4, &rc); if (!b) return rc;
        const float* rw = tensor(h, n + "/res_skip_" + li + "/W", (size_t)c.wn_c * n_rs * 4, &rc); if (!rw) return rc;
        const float* rb = tensor(h, n + "/res_skip_" + li + "/b", (size_t)n_rs * 4, &rc); if (!rb) return rc;
        mbexwn_op_t op = simple_conv(c.wn_k, c.wn_c, 2 * c.wn_c, d, (c.wn_causal ? (c.wn_k - 1) * d : (c.wn_k - 1) * d / 2));
        MBX_CUDA_CHECK(launch_conv1d(conv_args(c, op, c.steps_per_frame, rows, hbuf, w, b, nullptr, z), cx.g, cx.s));
        GateArgs ga{z, cond, act, rows, c.steps_per_frame, c.wn_c, c.wn_cond_lin_up, c.wn_gate};
        MBX_CUDA_CHECK(launch_gate(ga, cx.g, cx.s));
        mbexwn_op_t op2 = simple_conv(1, c.wn_c, n_rs, 1, 0);
        MBX_CUDA_CHECK(launch_conv1d(conv_args(c, op2, c.steps_per_frame, rows, act, rw, rb, nullptr, rs), cx.g, cx.s));
        ResSkipArgs ra{rs, hbuf, skip, rows, c.steps_per_frame, c.wn_c, n_rs, i == 0 ? 1 : 0};
        MBX_CUDA_CHECK(launch_resskip(ra, cx.g, cx.s));
        h->launches += 4;
    }
    return MBEXWN_OK;
}

static int forward_impl(mbexwn_handle_t h, const mbexwn_batch_t* b, int precision, void* workspace,
                        size_t workspace_bytes, cudaStream_t s) {
    const mbexwn_config_t& c = h->cfg;
    if (!b || b->n_frames <= 0 || b->n_utt <= 0) return fail(h, MBEXWN_ERR_INVALID, "empty batch");
    if (!b->frame_utt || !b->utt_begin || !b->utt_end || !b->chunk_first || !b->mel || !b->out)
        return fail(h, MBEXWN_ERR_INVALID, "batch has null pointers");
    if (precision < MBEXWN_PREC_FP32_SIMT || precision > MBEXWN_PREC_F16F8)
        return fail(h, MBEXWN_ERR_INVALID, "unknown precision");
    Ctx cx{h, b, FrameGrid{b->frame_utt, b->utt_begin, b->utt_end, b->n_frames, b->n_utt},
           carve(*h, b->n_frames, b->n_chunks, precision, h->debug_taps), reinterpret_cast<char*>(workspace), s};
    if (!workspace || workspace_bytes < cx.ws.total) return fail(h, MBEXWN_ERR_INVALID, "workspace too small");
    h->launches = 0;
    int rc = 0;
    const long long rows = (long long)b->n_frames * h->out_steps;      // rows of the sub-band signals (post net, PQMF)
    int stage = 0;
    h->ev_recorded = false;
    if (h->stage_timing && !h->ev_ready) {
        for (int i = 0; i <= MBEXWN_N_STAGES; ++i) MBX_CUDA_CHECK(cudaEventCreate(&h->ev[i]));
        h->ev_ready = true;
    }
    auto mark = [&]() { if (h->stage_timing) cudaEventRecord(h->ev[stage++], s); };
    mark();

    // (0) optional mel-derived RMS normaliser (NormMelComponents.normalize_inputs_by_rms, wavegen_1d.py:493-497, :638-769)
    const float* mel = b->mel;
    NormArgs na{};
    const float* norm_rms_prev = nullptr;
    if (c.norm_enable) {
        na.proj = tensor(h, "norm/proj", (size_t)c.mel_channels * (c.norm_proj_cols > 0 ? c.norm_proj_cols : 1) * 4, &rc); if (!na.proj) return rc;
        na.smooth_win = tensor(h, "norm/smooth_win", (size_t)c.norm_smooth_win * 4, &rc); if (!na.smooth_win) return rc;
        na.gwin = tensor(h, "norm/gwin", (size_t)c.norm_win * 4, &rc); if (!na.gwin) return rc;
        na.n_mel = c.mel_channels; na.proj_cols = c.norm_proj_cols; na.hop = c.hop; na.win = c.norm_win; na.ws = c.norm_smooth_win;
        na.off = c.norm_smooth_win / 2 + 2 * c.hop - c.norm_win / 2; na.iters = c.norm_iters; na.use_max_limit = c.norm_use_max_limit;
        na.proj_scale = c.norm_proj_scale; na.norm_fact = c.norm_fact; na.floor = c.norm_floor; na.compress_exp = c.norm_compress_exp;
        na.lin_scale = c.norm_lin_scale; na.lin_off = c.norm_lin_off; na.mel_scale = c.norm_mel_scale;
        MBX_CUDA_CHECK(launch_norm_mel(na, cx.g, b->mel, cx.p<float>("norm_rms_a"), cx.p<float>("norm_rms_b"),
                                       cx.p<float>("mel_norm"), &norm_rms_prev, &h->launches, s));
        mel = cx.p<float>("mel_norm");
    }

    // (1) F0 sub-net (generate_f0, custom_pulsed_generator.py:773-791)
    const float* f0 = b->f0_override;
    if (!f0) {
        rc = (precision == MBEXWN_PREC_FP32_SIMT || !h->tc_subnets)
                 ? run_subnet(cx, c.pp_ops, c.n_pp_ops, mel, cx.p<float>("F0"))
                 : run_subnet_tc(cx, c.pp_ops, c.n_pp_ops, mel, cx.p<float>("F0"));
        if (rc) return rc;
        f0 = cx.p<float>("F0");
    } else {
        MBX_CUDA_CHECK(cudaMemcpyAsync(cx.p<float>("F0"), f0, (size_t)b->n_frames * c.pulse_per_frame * 4,
                                       cudaMemcpyDeviceToDevice, s));
    }

    if (h->stop_after_f0) return MBEXWN_OK;
    mark();
    // (2) excitation generator head (generate_excitation, :886-906)
    {
        ExcitationArgs a{};
        a.f0 = f0; a.noise = b->noise; a.seed = b->seed; a.utt_ids = b->utt_ids;
        a.tables = tensor(h, "wavetable", (size_t)(c.wt_n_period + 1) * c.wt_n_tables * 4, &rc); if (!a.tables) return rc;
        a.n_period = c.wt_n_period; a.n_tables = c.wt_n_tables; a.pulse_rate = c.pulse_rate;
        a.nominal_f0 = c.wt_nominal_f0; a.min_tr = c.wt_min_transposition; a.max_tr = c.wt_max_transposition;
        a.grid_norm = c.wt_grid_norm; a.sigma = c.noise_sigma;
        a.pulse_per_frame = c.pulse_per_frame; a.steps_per_frame = c.steps_per_frame; a.pulse_channels = c.pulse_channels;
        a.subharm = c.wt_subharm;
        a.pqmf_taps = c.pulse_pqmf_taps;
        if (c.pulse_pqmf_taps > 0) {
            a.pqmf_ana = tensor(h, "pulse_pqmf", (size_t)c.pulse_channels * (c.pulse_pqmf_taps + 1) * 4, &rc); if (!a.pqmf_ana) return rc;
        }
        a.chunk = c.cumsum_chunk; a.cum = cx.p<float>("cum"); a.chunk_off = cx.p<float>("chunk_off");
        a.chunk_first = b->chunk_first; a.phase_carry = b->phase_carry; a.wn_in = cx.p<float>("wn_in"); a.ld_wn_in = c.wn_cin;
        a.phase_out = cx.p<float>("phase"); a.index_out = cx.p<int32_t>("index"); a.pulse_out = cx.p<float>("pulse");
        MBX_CUDA_CHECK(launch_excitation(a, cx.g, b->n_chunks, s));
        h->launches += 3 + (c.pulse_pqmf_taps > 0 ? 1 : 0);
    }

    mark();
    // (3) conditioning conv at mel rate; the x10 linear interpolation is fused into the gate (custom_AE_layers.py:282-289)
    auto run_cond = [&](const mbexwn_config_t& bc) -> int {
        const std::string n = std::string(bc.wn_name) + "/cond_";
        const int cout = 2 * bc.wn_c * bc.wn_cond_conv_up;
        const float* w = tensor(h, n + "/W", (size_t)bc.wn_cond_k * bc.mel_channels * cout * 4, &rc); if (!w) return rc;
        const float* bias = tensor(h, n + "/b", (size_t)cout * 4, &rc); if (!bias) return rc;
        mbexwn_op_t op = simple_conv(bc.wn_cond_k, bc.mel_channels, cout, 1, (bc.wn_causal ? bc.wn_cond_k - 1 : (bc.wn_cond_k - 1) / 2));
        if (precision != MBEXWN_PREC_FP32_SIMT && h->tc_subnets && tc_eligible(op)) {
            const int cin_pad = round64(bc.mel_channels);
            const float* wt = tensor(h, n + "/tc/W", (size_t)cout * 2 * op.k * cin_pad * 2, &rc); if (!wt) return rc;
            MBX_RC(wn_tc_pack(mel, cx.p<char>("mel_hl"), b->n_frames, bc.mel_channels, cin_pad, 1, 0, 0, PAD_ZERO, cx.g, s, &h->error));
            TcConvArgs a{};
            a.a_hilo = cx.p<char>("mel_hl"); a.rows = b->n_frames; a.cin_pad = cin_pad; a.w = wt; a.cout = cout; a.k = op.k;
            a.dilation = 1; a.pad_l = op.pad_l; a.bias = bias; a.act = ACT_NONE; a.rate = 1;
            a.out_f32 = cx.p<float>("cond"); a.ld_out = cout;
            MBX_RC(wn_tc_conv(h->tc, a, cx.g, s, &h->error));
            h->launches += 2;
        } else {
            MBX_CUDA_CHECK(launch_conv1d(conv_args(bc, op, 1, b->n_frames, mel, w, bias, nullptr, cx.p<float>("cond")), cx.g, s));
            h->launches++;
        }
        return MBEXWN_OK;
    };
    if ((rc = run_cond(h->blocks[0].cfg))) return rc;

    mark();
    // (4) WaveNet blocks (custom_pulsed_generator.py:908-910): WaveNetAE.call (custom_AE_layers.py:273-346), then the block's
    //     sub-pixel up-sampling conv (:519-526, :571-573) when its factor is > 1.  The released models are one block without it.
    const int out_pad = wn_tc_out_pad(c);
    const float* post_in = cx.p<float>("wn_out");
    for (size_t ib = 0; ib < h->blocks.size(); ++ib) {
        const WnBlock& blk = h->blocks[ib];
        const mbexwn_config_t& bc = blk.cfg;
        const bool last_block = ib + 1 == h->blocks.size();
        const long long brows = (long long)b->n_frames * bc.steps_per_frame;
        // block 0 reads the excitation head's rows; a later block reads what the previous one left: its up-sampling conv's output
        // (blk_in, wn_cout floats per row) or, without up-sampling, the previous WaveNet output in place (row pitch out_pad)
        const bool prev_up = ib > 0 && h->blocks[ib - 1].up > 1;
        const float* wn_in = ib == 0 ? cx.p<float>("wn_in") : (prev_up ? cx.p<float>("blk_in") : cx.p<float>("wn_out"));
        const int ld_in = ib == 0 ? bc.wn_cin : (prev_up ? bc.wn_cin : out_pad);
        if (ib > 0 && (rc = run_cond(bc))) return rc;         // every block has its own conditioning conv
        if (precision == MBEXWN_PREC_FP32_SIMT) {
            rc = wavenet_fp32(cx, bc, wn_in, ld_in);
            if (rc) return rc;
            // `end` 1x1 over the skip sum (custom_AE_layers.py:340); the tensor-core path folds it into the res_skip matrices
            const std::string n = std::string(bc.wn_name) + "/end";
            const float* w = tensor(h, n + "/W", (size_t)bc.wn_c * bc.wn_cout * 4, &rc); if (!w) return rc;
            const float* bias = tensor(h, n + "/b", (size_t)bc.wn_cout * 4, &rc); if (!bias) return rc;
            mbexwn_op_t op = simple_conv(1, bc.wn_c, bc.wn_cout, 1, 0);
            ConvArgs a = conv_args(bc, op, bc.steps_per_frame, brows, cx.p<float>("skip"), w, bias, nullptr, cx.p<float>("wn_out"));
            a.ld_out = out_pad;
            // the row pitch is padded to 32 channels and read back as float4s: define the padding columns
            if (out_pad != bc.wn_cout) MBX_CUDA_CHECK(cudaMemsetAsync(cx.p<float>("wn_out"), 0, (size_t)brows * out_pad * 4, s));
            MBX_CUDA_CHECK(launch_conv1d(a, cx.g, s));
            h->launches++;
        } else {
            rc = wn_tc_forward(h->tc, bc, cx.g, precision, wn_in, ld_in, cx.p<float>("cond"), cx.p<float>("wn_out"),
                               [&](const char* nm) { return (void*)cx.p<char>(nm); },
                               [&](const std::string& nm, size_t bytes) { int r = 0; return (const void*)tensor(h, nm, bytes, &r); },
                               s, &h->launches, &h->error);
            if (rc) return rc;
        }
        if (blk.up > 1) {
            // TF2C_Conv1DUpDownSample (conv_layers.py:250-255): k = 3 conv to wn_cout x up channels, channel c' of row t becomes
            // row t x up + c' / wn_cout, channel c' % wn_cout
            const int cout = bc.wn_cout * blk.up;
            const float* w = tensor(h, blk.up_name + "/W", (size_t)3 * bc.wn_cout * cout * 4, &rc); if (!w) return rc;
            const float* bias = tensor(h, blk.up_name + "/b", (size_t)cout * 4, &rc); if (!bias) return rc;
            mbexwn_op_t op = simple_conv(3, bc.wn_cout, cout, 1, bc.wn_causal ? 2 : 1);
            ConvArgs a = conv_args(bc, op, bc.steps_per_frame, brows, cx.p<float>("wn_out"), w, bias, nullptr, cx.p<float>("blk_in"));
            a.ld_x = out_pad;
            a.sub_ch = bc.wn_cout;
            a.ld_out = last_block ? out_pad : bc.wn_cout;
            if (last_block && out_pad != bc.wn_cout)
                MBX_CUDA_CHECK(cudaMemsetAsync(cx.p<float>("blk_in"), 0, (size_t)brows * blk.up * out_pad * 4, s));
            MBX_CUDA_CHECK(launch_conv1d(a, cx.g, s));
            h->launches++;
            if (last_block) post_in = cx.p<float>("blk_in");
        }
    }
    mark();

    // (5) post 1x1 over the WaveNet output (custom_pulsed_generator.py:913-914), then PQMF synthesis (:920-921).
    //     Tensor-core path: `end` (custom_AE_layers.py:340) is already folded into the res_skip matrices; the fp32
    //     variant applies it here to the skip sum.
    {
        const std::string pn = c.post_name;
        const float* w = tensor(h, pn + "/W", (size_t)c.wn_cout * c.subbands * 4, &rc); if (!w) return rc;
        const float* bias = tensor(h, pn + "/b", (size_t)c.subbands * 4, &rc); if (!bias) return rc;
        const float* poly = tensor(h, "pqmf_poly", (size_t)c.pqmf_q * c.subbands * c.subbands * 4, &rc); if (!poly) return rc;
        PostPqmfArgs fa{};
        fa.wn_out = post_in; fa.ld = out_pad; fa.cin = c.wn_cout; fa.post_w = w; fa.post_b = bias; fa.poly = poly;
        fa.sub_out = h->debug_taps ? cx.p<float>("subbands") : nullptr; fa.out = cx.p<float>("excitation");
        fa.rows = rows; fa.steps_per_frame = h->out_steps; fa.S = c.subbands; fa.Q = c.pqmf_q; fa.back = c.pqmf_back;
        if (c.ps_mode != 0) {
            // no STFT-domain filter: the PQMF output is the signal (custom_pulsed_generator.py:666-674)
            fa.out = b->out;
            if (!post_pqmf_supported(fa)) return fail(h, MBEXWN_ERR_UNSUPPORTED, "ps_use_stft = False / ps_off need the fused post + PQMF kernel");
            if (c.ps_mode == 1) {
                // per-band log gains from the PS sub-net (generate_multiband_gain, :857-884), applied inside the fused kernel
                rc = (precision == MBEXWN_PREC_FP32_SIMT || !h->tc_subnets)
                         ? run_subnet(cx, c.ps_ops, c.n_ps_ops, mel, cx.p<float>("ceps"))
                         : run_subnet_tc(cx, c.ps_ops, c.n_ps_ops, mel, cx.p<float>("ceps"));
                if (rc) return rc;
                fa.log_gain = cx.p<float>("ceps"); fa.gain_up = c.hop; fa.gain_center = c.ps_preserve_energy;
            }
        }
        if (post_pqmf_supported(fa)) {
            MBX_CUDA_CHECK(launch_post_pqmf(fa, cx.g, s));
            h->launches += 1;
        } else {
            mbexwn_op_t op = simple_conv(1, c.wn_cout, c.subbands, 1, 0);
            ConvArgs a = conv_args(c, op, h->out_steps, rows, post_in, w, bias, nullptr, cx.p<float>("subbands"));
            a.ld_x = out_pad;
            MBX_CUDA_CHECK(launch_conv1d(a, cx.g, s));
            PqmfArgs pa{};
            pa.sub = cx.p<float>("subbands");
            pa.poly = poly;
            pa.out = cx.p<float>("excitation"); pa.rows = rows; pa.steps_per_frame = h->out_steps;
            pa.S = c.subbands; pa.Q = c.pqmf_q; pa.back = c.pqmf_back;
            MBX_CUDA_CHECK(launch_pqmf(pa, cx.g, s));
            h->launches += 2;
        }
    }

    mark();
    if (c.ps_mode != 0) {
        mark();
        if (c.norm_enable) {
            float* tap = cx.ws.slots.count("norm_gain") ? cx.p<float>("norm_gain") : nullptr;
            MBX_CUDA_CHECK(launch_norm_apply(na, cx.g, norm_rms_prev, b->out, tap, c.hop, s));
            h->launches += 1;
        }
        mark();
        h->ev_recorded = h->stage_timing != 0;
        return MBEXWN_OK;
    }
    // (6) VTF sub-net -> cepstrum (generate_specenv, :793-799)
    rc = (precision == MBEXWN_PREC_FP32_SIMT || !h->tc_subnets)
             ? run_subnet(cx, c.ps_ops, c.n_ps_ops, mel, cx.p<float>("ceps"))
             : run_subnet_tc(cx, c.ps_ops, c.n_ps_ops, mel, cx.p<float>("ceps"));
    if (rc) return rc;
    mark();

    // (7) STFT-domain filtering + overlap-add (:681-724, :801-836)
    {
        StftFilterArgs a{};
        a.exc = cx.p<float>("excitation"); a.ceps = cx.p<float>("ceps"); a.f0 = f0;
        if (c.n_lifters > 0) {
            a.lifters = tensor(h, "lifters", (size_t)c.n_lifters * c.n_ceps * 4, &rc); if (!a.lifters) return rc;
            a.lifter_grid = tensor(h, "lifter_grid", (size_t)c.n_lifters * 4, &rc); if (!a.lifter_grid) return rc;
            a.f0_smooth = tensor(h, "f0_smooth", (size_t)c.n_smooth * 4, &rc); if (!a.f0_smooth) return rc;
        }
        a.n_lift = c.n_lifters; a.n_smooth = c.n_smooth; a.pulse_per_frame = c.pulse_per_frame;
        a.window = tensor(h, "window", (size_t)c.stft_win * 4, &rc); if (!a.window) return rc;
        a.inv_window = tensor(h, "inv_window", (size_t)c.stft_win * 4, &rc); if (!a.inv_window) return rc;
        a.twiddle = reinterpret_cast<const float2*>(tensor(h, "twiddle", (size_t)c.fft_size * 4, &rc)); if (!a.twiddle) return rc;
        a.frames_out = cx.p<float>("frames"); a.vtf_out = cx.p<float>("vtf"); a.lifter_index_out = cx.p<int32_t>("lifter_index");
        a.n_frames = b->n_frames; a.hop = c.hop; a.win = c.stft_win; a.fft = c.fft_size; a.n_ceps = c.n_ceps;
        a.max_log_range = c.filter_max_log_range;
        MBX_CUDA_CHECK(launch_stft_filter(a, cx.g, s));
        OlaArgs oa{cx.p<float>("frames"), b->out, b->n_frames, c.hop, c.stft_win};
        MBX_CUDA_CHECK(launch_ola(oa, cx.g, s));
        h->launches += 2;
        if (c.norm_enable) {                                  // signal * upsampled_rms (wavegen_1d.py:504-507)
            float* tap = cx.ws.slots.count("norm_gain") ? cx.p<float>("norm_gain") : nullptr;
            MBX_CUDA_CHECK(launch_norm_apply(na, cx.g, norm_rms_prev, b->out, tap, c.hop, s));
            h->launches += 1;
        }
    }
    mark();
    h->ev_recorded = h->stage_timing != 0;
    return MBEXWN_OK;
}

}  // namespace mbx

// ---- extern "C" ------------------------------------------------------------------------------------------

extern "C" {

int mbexwn_abi_version(void) { return MBEXWN_ABI_VERSION; }

int mbexwn_create(const mbexwn_config_t* cfg, mbexwn_handle_t* out) {
    if (!cfg || !out) return MBEXWN_ERR_INVALID;
    *out = nullptr;
    if (cfg->abi_version != MBEXWN_ABI_VERSION) return MBEXWN_ERR_INVALID;
    if (cfg->wn_layers < 1 || cfg->wn_layers > MBEXWN_MAX_LAYERS || cfg->n_pp_ops > MBEXWN_MAX_OPS ||
        cfg->n_ps_ops > MBEXWN_MAX_OPS || cfg->n_pp_ops < 1 || (cfg->n_ps_ops < 1 && cfg->ps_mode != 2) ||
        cfg->ps_mode < 0 || cfg->ps_mode > 2)
        return MBEXWN_ERR_INVALID;
    if (cfg->ps_mode == 1 && cfg->ps_ops[cfg->n_ps_ops - 1].ch_out != cfg->subbands) return MBEXWN_ERR_INVALID;
    if (cfg->wn_n_blocks < 0 || cfg->wn_n_blocks > MBEXWN_MAX_BLOCKS) return MBEXWN_ERR_INVALID;
    int out_steps = cfg->steps_per_frame;
    for (int i = 0; i < cfg->wn_n_blocks; ++i) {
        const mbexwn_wn_block_t& bl = cfg->wn_blocks[i];
        if (bl.c < 2 || bl.cond_conv_up < 1 || bl.up < 1 || !bl.name[0] || (bl.up > 1 && !bl.up_name[0])) return MBEXWN_ERR_INVALID;
        if (out_steps != bl.cond_conv_up * cfg->wn_cond_lin_up) return MBEXWN_ERR_INVALID;   // custom_pulsed_generator.py:469
        out_steps *= bl.up;
    }
    if (out_steps * cfg->subbands != cfg->hop) return MBEXWN_ERR_INVALID;
    if (cfg->steps_per_frame * cfg->pulse_channels != cfg->pulse_per_frame) return MBEXWN_ERR_INVALID;
    if (cfg->wt_subharm < 0 || cfg->wn_cin != cfg->pulse_channels * (1 + cfg->wt_subharm) + (cfg->noise_sigma != 0.f ? 1 : 0))
        return MBEXWN_ERR_INVALID;
    if (cfg->pulse_pqmf_taps < 0 || (cfg->pulse_pqmf_taps & 1)) return MBEXWN_ERR_INVALID;
    if (cfg->fft_size < cfg->stft_win || (cfg->fft_size & (cfg->fft_size - 1))) return MBEXWN_ERR_INVALID;
    if (cfg->stft_win != 4 * cfg->hop) return MBEXWN_ERR_UNSUPPORTED;     // 4x overlap (wavegen_1d.py:592)
    if (cfg->norm_enable && (cfg->norm_iters < 1 || cfg->norm_win != 4 * cfg->hop || cfg->norm_smooth_win < 1 ||
                             cfg->mel_channels > 256 || cfg->norm_fact <= 0.f))
        return MBEXWN_ERR_UNSUPPORTED;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return MBEXWN_ERR_CUDA;                          // no CPU fallback
    mbexwn_handle_s* h = new mbexwn_handle_s();
    h->cfg = *cfg;
    h->device = dev;
    h->out_steps = out_steps;
    if (cfg->wn_n_blocks == 0) {
        h->blocks.push_back(WnBlock{*cfg, 1, std::string()});
    } else {
        int rate = cfg->steps_per_frame;
        for (int i = 0; i < cfg->wn_n_blocks; ++i) {
            const mbexwn_wn_block_t& bl = cfg->wn_blocks[i];
            WnBlock wb{*cfg, bl.up, std::string(bl.up_name)};
            wb.cfg.wn_c = bl.c;
            wb.cfg.wn_cond_conv_up = bl.cond_conv_up;
            wb.cfg.steps_per_frame = rate;
            if (i > 0) wb.cfg.wn_cin = cfg->wn_cout;
            std::memcpy(wb.cfg.wn_name, bl.name, sizeof(wb.cfg.wn_name));
            h->blocks.push_back(wb);
            rate *= bl.up;
        }
    }
    // sticky range-guard word of the f16f8 path in mapped host memory: the kernels write it, the host reads it after a
    // synchronisation without any copy
    if (cudaHostAlloc(reinterpret_cast<void**>(&h->range_flag), 64, cudaHostAllocMapped) == cudaSuccess && h->range_flag) {
        *h->range_flag = 0;
        int* dptr = nullptr;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), h->range_flag, 0) == cudaSuccess) h->tc.range_flag = dptr;
    }
    *out = h;
    return MBEXWN_OK;
}

void mbexwn_destroy(mbexwn_handle_t h) {
    if (!h) return;
    if (h->ev_ready) for (int i = 0; i <= MBEXWN_N_STAGES; ++i) cudaEventDestroy(h->ev[i]);
    if (h->pipe_ready) {
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(h->ev_h2d[i]); cudaEventDestroy(h->ev_done[i]); cudaEventDestroy(h->ev_d2h[i]); }
        cudaStreamDestroy(h->copy_in);
        cudaStreamDestroy(h->copy_out);
    }
    mbx::wn_tc_destroy(h->tc);
    if (h->range_flag) cudaFreeHost(h->range_flag);
    delete h;
}

const char* mbexwn_last_error(mbexwn_handle_t h) { return h ? h->error.c_str() : "null handle"; }

int mbexwn_set_tensor(mbexwn_handle_t h, const char* name, const void* dev_ptr, size_t n_bytes) {
    if (!h || !name || !dev_ptr) return MBEXWN_ERR_INVALID;
    h->tensors[name] = std::make_pair(dev_ptr, n_bytes);
    mbx::wn_tc_invalidate(h->tc);
    return MBEXWN_OK;
}

int mbexwn_set_scalar(mbexwn_handle_t h, const char* name, float value) {
    if (!h || !name) return MBEXWN_ERR_INVALID;
    h->host_scalars[name] = value;
    return MBEXWN_OK;
}

size_t mbexwn_workspace_bytes(mbexwn_handle_t h, int32_t n_frames, int32_t n_chunks, int32_t precision) {
    if (!h || n_frames <= 0) return 0;
    return mbx::carve(*h, n_frames, n_chunks, precision, h->debug_taps).total;
}

int mbexwn_forward(mbexwn_handle_t h, const mbexwn_batch_t* batch, int32_t precision, void* workspace,
                   size_t workspace_bytes, void* cuda_stream) {
    if (!h) return MBEXWN_ERR_INVALID;
    return mbx::forward_impl(h, batch, precision, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(cuda_stream));
}

int mbexwn_forward_host(mbexwn_handle_t h, const mbexwn_batch_t* batch, int32_t precision, const float* mel_host,
                        const float* noise_host, float* out_host, void* workspace, size_t workspace_bytes,
                        void* cuda_stream) {
    if (!h || !batch || !mel_host || !out_host) return MBEXWN_ERR_INVALID;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    const mbexwn_config_t& c = h->cfg;
    MBX_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(batch->mel), mel_host,
                                   (size_t)batch->n_frames * c.mel_channels * 4, cudaMemcpyHostToDevice, s));
    if (noise_host && batch->noise)
        MBX_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(batch->noise), noise_host,
                                       (size_t)batch->n_frames * c.steps_per_frame * 4, cudaMemcpyHostToDevice, s));
    int rc = mbx::forward_impl(h, batch, precision, workspace, workspace_bytes, s);
    if (rc) return rc;
    MBX_CUDA_CHECK(cudaMemcpyAsync(out_host, batch->out, (size_t)batch->n_frames * c.hop * 4, cudaMemcpyDeviceToHost, s));
    MBX_CUDA_CHECK(cudaStreamSynchronize(s));
    return MBEXWN_OK;
}

int mbexwn_forward_host_begin(mbexwn_handle_t h, int32_t slot, const mbexwn_batch_t* batch, int32_t precision,
                              const float* mel_host, const float* noise_host, float* out_host, void* workspace,
                              size_t workspace_bytes, void* cuda_stream) {
    if (!h || !batch || !mel_host || !out_host || slot < 0 || slot > 1) return MBEXWN_ERR_INVALID;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    const mbexwn_config_t& c = h->cfg;
    if (!h->pipe_ready) {
        MBX_CUDA_CHECK(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
        MBX_CUDA_CHECK(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            MBX_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
            MBX_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            MBX_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
        }
        h->pipe_ready = true;
    }
    // H2D of this call's inputs: the slot's device input buffers were last read by the forward two calls ago
    if (h->slot_busy[slot]) MBX_CUDA_CHECK(cudaStreamWaitEvent(h->copy_in, h->ev_done[slot], 0));
    MBX_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(batch->mel), mel_host, (size_t)batch->n_frames * c.mel_channels * 4,
                                   cudaMemcpyHostToDevice, h->copy_in));
    if (noise_host && batch->noise)
        MBX_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(batch->noise), noise_host,
                                       (size_t)batch->n_frames * c.steps_per_frame * 4, cudaMemcpyHostToDevice, h->copy_in));
    MBX_CUDA_CHECK(cudaEventRecord(h->ev_h2d[slot], h->copy_in));
    // forward on the caller's stream: needs the inputs, and the slot's output buffer drained by the D2H two calls ago
    MBX_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_h2d[slot], 0));
    if (h->slot_busy[slot]) MBX_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_d2h[slot], 0));
    int rc = mbx::forward_impl(h, batch, precision, workspace, workspace_bytes, s);
    if (rc) return rc;
    MBX_CUDA_CHECK(cudaEventRecord(h->ev_done[slot], s));
    // D2H of the waveform under the next call's kernels
    MBX_CUDA_CHECK(cudaStreamWaitEvent(h->copy_out, h->ev_done[slot], 0));
    MBX_CUDA_CHECK(cudaMemcpyAsync(out_host, batch->out, (size_t)batch->n_frames * c.hop * 4, cudaMemcpyDeviceToHost, h->copy_out));
    MBX_CUDA_CHECK(cudaEventRecord(h->ev_d2h[slot], h->copy_out));
    h->slot_busy[slot] = true;
    return MBEXWN_OK;
}

int mbexwn_forward_host_wait(mbexwn_handle_t h, int32_t slot) {
    if (!h || slot < 0 || slot > 1) return MBEXWN_ERR_INVALID;
    if (!h->pipe_ready || !h->slot_busy[slot]) return MBEXWN_OK;
    MBX_CUDA_CHECK(cudaEventSynchronize(h->ev_d2h[slot]));
    return MBEXWN_OK;
}

int mbexwn_tap(mbexwn_handle_t h, const char* name, int32_t n_frames, int32_t n_chunks, int32_t precision,
               size_t* offset_bytes, size_t* n_bytes) {
    if (!h || !name || !offset_bytes || !n_bytes) return MBEXWN_ERR_INVALID;
    mbx::Workspace w = mbx::carve(*h, n_frames, n_chunks, precision, h->debug_taps);
    auto it = w.slots.find(name);
    if (it == w.slots.end()) return mbx::fail(h, MBEXWN_ERR_MISSING, std::string("unknown tap: ") + name);
    *offset_bytes = it->second.off;
    *n_bytes = it->second.bytes;
    return MBEXWN_OK;
}

int mbexwn_last_launch_count(mbexwn_handle_t h) { return h ? h->launches : 0; }

int mbexwn_get_info(mbexwn_handle_t h, const char* name, int32_t* value) {
    if (!h || !name || !value) return MBEXWN_ERR_INVALID;
    if (!strcmp(name, "tc_last_fused")) { *value = h->tc.last_fused; return MBEXWN_OK; }
    if (!strcmp(name, "tc_last_cluster")) { *value = h->tc.last_cluster; return MBEXWN_OK; }
    if (!strcmp(name, "tc_max_quads")) { *value = h->tc.last_quads; return MBEXWN_OK; }
    h->error = std::string("unknown info: ") + name;
    return MBEXWN_ERR_INVALID;
}

int mbexwn_set_option(mbexwn_handle_t h, const char* name, int32_t value) {
    if (!h || !name) return MBEXWN_ERR_INVALID;
    if (!strcmp(name, "debug_taps")) { h->debug_taps = value ? 1 : 0; return MBEXWN_OK; }
    if (!strcmp(name, "stage_timing")) { h->stage_timing = value ? 1 : 0; h->tc.time_launches = h->stage_timing; return MBEXWN_OK; }
    if (!strcmp(name, "tc_cta_group")) { h->tc.cta_group = value == 2 ? 2 : 1; return MBEXWN_OK; }
    if (!strcmp(name, "tc_subnets")) { h->tc_subnets = value ? 1 : 0; return MBEXWN_OK; }
    if (!strcmp(name, "fuse_tail")) { h->fuse_tail = value ? 1 : 0; return MBEXWN_OK; }
    if (!strcmp(name, "tc_cond_stage")) { h->tc.cond_stage = value ? 1 : 0; return MBEXWN_OK; }
    if (!strcmp(name, "stop_after_f0")) { h->stop_after_f0 = value ? 1 : 0; return MBEXWN_OK; }
    if (!strcmp(name, "tc_debug")) { h->tc.debug = value; return MBEXWN_OK; }
    if (!strcmp(name, "tc_fused")) { h->tc.fused = value < 0 ? 0 : (value > 2 ? 2 : value); return MBEXWN_OK; }
    if (!strcmp(name, "tc_slab")) { h->tc.slab = value ? 1 : 0; return MBEXWN_OK; }
    if (!strcmp(name, "tc_ring_a")) { h->tc.n_a = value; return MBEXWN_OK; }
    if (!strcmp(name, "tc_interleave")) { h->tc.interleave = value ? 1 : 0; return MBEXWN_OK; }
    if (!strcmp(name, "tc_discard")) { h->tc.discard = value ? 1 : 0; return MBEXWN_OK; }
    if (!strcmp(name, "tc_cluster")) { h->tc.cluster = value == 4 ? 4 : 2; return MBEXWN_OK; }
    if (!strcmp(name, "tc_trace")) { h->tc.trace_on = value; return MBEXWN_OK; }
    if (!strcmp(name, "tc8_h_lo")) { h->tc.sh_h_lo = value; return MBEXWN_OK; }
    if (!strcmp(name, "tc8_a_lo")) { h->tc.sh_a_lo = value; return MBEXWN_OK; }
    return mbx::fail(h, MBEXWN_ERR_INVALID, std::string("unknown option: ") + name);
}

int mbexwn_stage_ms(mbexwn_handle_t h, float* ms) {
    if (!h || !ms) return MBEXWN_ERR_INVALID;
    if (!h->ev_recorded) return mbx::fail(h, MBEXWN_ERR_INVALID, "no stage timing recorded (set option stage_timing)");
    MBX_CUDA_CHECK(cudaEventSynchronize(h->ev[MBEXWN_N_STAGES]));
    for (int i = 0; i < MBEXWN_N_STAGES; ++i) MBX_CUDA_CHECK(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
    return MBEXWN_OK;
}

int mbexwn_wavenet_launch_ms(mbexwn_handle_t h, float* gate_ms, float* resskip_ms, int32_t* n_layers) {
    if (!h || !gate_ms || !resskip_ms || !n_layers) return MBEXWN_ERR_INVALID;
    int n = 0;
    int rc = mbx::wn_tc_launch_ms(h->tc, gate_ms, resskip_ms, &n);
    if (rc) return mbx::fail(h, rc, "no WaveNet launch timing recorded (set option stage_timing, tensor-core precision)");
    *n_layers = n;
    return MBEXWN_OK;
}

int mbexwn_phase_carry(mbexwn_handle_t h, const float* f0_dev, int64_t n_samples, float* run_out_dev, void* cuda_stream) {
    if (!h || !f0_dev || !run_out_dev || n_samples < 1) return MBEXWN_ERR_INVALID;
    MBX_CUDA_CHECK(mbx::launch_phase_carry(f0_dev, n_samples, h->cfg.pulse_rate, h->cfg.cumsum_chunk, run_out_dev,
                                           reinterpret_cast<cudaStream_t>(cuda_stream)));
    return MBEXWN_OK;
}

int mbexwn_gather_rows(mbexwn_handle_t h, const float* src_dev, float* dst_dev, int32_t row_elems, const int64_t* seg_dev,
                       int32_t n_seg, int32_t max_rows, void* cuda_stream) {
    if (!h || !src_dev || !dst_dev || !seg_dev) return MBEXWN_ERR_INVALID;
    MBX_CUDA_CHECK(mbx::launch_gather_rows(src_dev, dst_dev, row_elems, reinterpret_cast<const long long*>(seg_dev), n_seg, max_rows,
                                           reinterpret_cast<cudaStream_t>(cuda_stream)));
    return MBEXWN_OK;
}

int mbexwn_range_status(mbexwn_handle_t h, int32_t* flags, int32_t reset) {
    if (!h || !flags) return MBEXWN_ERR_INVALID;
    *flags = h->range_flag ? *reinterpret_cast<volatile int*>(h->range_flag) : 0;
    if (reset && h->range_flag) *reinterpret_cast<volatile int*>(h->range_flag) = 0;
    return MBEXWN_OK;
}

int64_t mbexwn_tc_trace_read(mbexwn_handle_t h, uint32_t* out, int64_t n_words) {
    if (!h || !out || n_words <= 0) return -1;
    return mbx::wn_tc_read_trace(h->tc, out, n_words);
}

int mbexwn_k_conv1d(mbexwn_handle_t h, const mbexwn_batch_t* b, const mbexwn_op_t* op, int32_t rate, const float* x,
                    const float* w, const float* bias, const float* alpha, float* out, void* cuda_stream) {
    if (!h || !b || !op || !x || !w || !bias || !out) return MBEXWN_ERR_INVALID;
    mbx::FrameGrid g{b->frame_utt, b->utt_begin, b->utt_end, b->n_frames, b->n_utt};
    mbx::ConvArgs a = mbx::conv_args(h->cfg, *op, rate, (long long)b->n_frames * rate, x, w, bias, alpha, out);
    MBX_CUDA_CHECK(mbx::launch_conv1d(a, g, reinterpret_cast<cudaStream_t>(cuda_stream)));
    return MBEXWN_OK;
}

int mbexwn_k_lininterp(mbexwn_handle_t h, const mbexwn_batch_t* b, const mbexwn_op_t* op, const float* x,
                       const float* alpha, float* out, void* cuda_stream) {
    if (!h || !b || !op || !x || !out) return MBEXWN_ERR_INVALID;
    mbx::FrameGrid g{b->frame_utt, b->utt_begin, b->utt_end, b->n_frames, b->n_utt};
    mbx::LinInterpArgs a{};
    a.x = x; a.out = out; a.rows_in = (long long)b->n_frames * op->rate_in; a.rate_in = op->rate_in; a.ch = op->ch_out;
    a.up = op->up; a.act = op->act; a.alpha = alpha; a.leaky = h->cfg.leaky_alpha;
    a.a0 = h->cfg.f0_span; a.a1 = h->cfg.f0_min;
    MBX_CUDA_CHECK(mbx::launch_lininterp(a, g, reinterpret_cast<cudaStream_t>(cuda_stream)));
    return MBEXWN_OK;
}

int mbexwn_k_tc_gemm(mbexwn_handle_t h, const void* a_bf16, int64_t rows, int32_t a_cols, const void* b_bf16, int32_t n,
                     int32_t b_cols, const int32_t* kblocks, int32_t n_kb, float* out, void* cuda_stream) {
    if (!h || !a_bf16 || !b_bf16 || !kblocks || !out) return MBEXWN_ERR_INVALID;
    return mbx::wn_tc_gemm_test(h->tc, a_bf16, rows, a_cols, b_bf16, n, b_cols, kblocks, n_kb, out,
                                reinterpret_cast<cudaStream_t>(cuda_stream), &h->error);
}

int mbexwn_k_tc_gemm_f16f8(mbexwn_handle_t h, const void* a, int64_t rows, int32_t a_cpad, const void* b, int32_t n,
                           int32_t b_k, const int32_t* kblocks, int32_t n_kb, float* out, void* cuda_stream) {
    if (!h || !a || !b || !kblocks || !out) return MBEXWN_ERR_INVALID;
    return mbx::wn_tc_gemm_test_f16f8(h->tc, a, rows, a_cpad, b, n, b_k, kblocks, n_kb, out,
                                      reinterpret_cast<cudaStream_t>(cuda_stream), &h->error);
}

static thread_local std::string g_error;

const char* mbexwn_global_error(void) { return g_error.c_str(); }

static int analysis_fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

int mbexwn_mel_analysis(const mbexwn_analysis_config_t* c, const mbexwn_analysis_batch_t* b, void* cuda_stream) {
    if (!c || !b) return analysis_fail(MBEXWN_ERR_INVALID, "null config / batch");
    if (!c->window || !c->twiddle || !c->mel_lo || !c->mel_cnt || !c->mel_off || !c->mel_w)
        return analysis_fail(MBEXWN_ERR_INVALID, "analysis config: a table pointer is null");
    if (b->n_utt < 0 || b->n_pairs < 0) return analysis_fail(MBEXWN_ERR_INVALID, "negative batch size");
    if (b->n_utt == 0 || b->n_pairs == 0) return MBEXWN_OK;
    if (!b->sample_begin || !b->n_samples || !b->frame_begin || !b->pair_first || !b->audio || !b->mel)
        return analysis_fail(MBEXWN_ERR_INVALID, "analysis batch: a buffer pointer is null");
    mbx::MelAnalysisArgs a{};
    a.audio = b->audio; a.sample_begin = reinterpret_cast<const long long*>(b->sample_begin); a.n_samples = b->n_samples;
    a.frame_begin = b->frame_begin; a.pair_first = b->pair_first; a.n_utt = b->n_utt; a.n_pairs = b->n_pairs;
    a.window = c->window; a.twiddle = reinterpret_cast<const float2*>(c->twiddle);
    a.mel_lo = c->mel_lo; a.mel_cnt = c->mel_cnt; a.mel_off = c->mel_off; a.mel_w = c->mel_w;
    a.hop = c->hop; a.win = c->win; a.fft = c->fft_size; a.n_mel = c->n_mel;
    a.mode = c->mode; a.lin_scale = c->lin_scale; a.lin_off = c->lin_off; a.log_scale = c->log_scale; a.floor = c->floor;
    a.mel_out = b->mel; a.mag_out = b->mag_tap;
    if (c->mode < 0 || c->mode > 2) return analysis_fail(MBEXWN_ERR_INVALID, "analysis config: mode must be 0, 1 or 2");
    if (!mbx::mel_analysis_supported(a))
        return analysis_fail(MBEXWN_ERR_UNSUPPORTED, "mel analysis is built for fft_size 2048, win <= 2048, n_mel <= 128");
    cudaError_t e = mbx::launch_mel_analysis(a, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (e != cudaSuccess) return analysis_fail(MBEXWN_ERR_CUDA, std::string("mel_analysis2048_kernel: ") + cudaGetErrorString(e));
    return MBEXWN_OK;
}

int mbexwn_mel_analysis_host(const mbexwn_analysis_config_t* c, const mbexwn_analysis_batch_t* b, const float* audio_host,
                             float* mel_host, void* cuda_stream) {
    if (!c || !b || !audio_host || !mel_host) return analysis_fail(MBEXWN_ERR_INVALID, "null argument");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    cudaError_t e = cudaMemcpyAsync(const_cast<float*>(b->audio), audio_host, (size_t)b->n_samples_total * 4,
                                    cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return analysis_fail(MBEXWN_ERR_CUDA, std::string("H2D audio: ") + cudaGetErrorString(e));
    int rc = mbexwn_mel_analysis(c, b, cuda_stream);
    if (rc) return rc;
    e = cudaMemcpyAsync(mel_host, b->mel, (size_t)b->n_frames * c->n_mel * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return analysis_fail(MBEXWN_ERR_CUDA, std::string("D2H mel: ") + cudaGetErrorString(e));
    return MBEXWN_OK;
}

}  // extern "C"
