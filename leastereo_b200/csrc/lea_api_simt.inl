// C-ABI entry points of the CUDA-core kernels (argument validation + launch).  Included by leastereo_b200.cu
// (nvcc, the product) and by tests/emu/emu_lib.cpp (g++, CPU emulation for the no-GPU unit tests).

static int lea_check_vol(const lea_vol* v, const char* what) {
    LEA_CHECK(v != nullptr && v->data != nullptr, "%s: null volume", what);
    LEA_CHECK(v->B > 0 && v->D > 0 && v->H > 0 && v->W > 0, "%s: empty volume", what);
    LEA_CHECK(v->C > 0 && (v->C & 7) == 0, "%s: channel count %d is not a multiple of 8", what, v->C);
    LEA_CHECK(v->P >= 1 && v->P <= 3, "%s: planes must be 1..3 (got %d)", what, v->P);
    LEA_CHECK((((uintptr_t)v->data) & 15) == 0, "%s: data pointer is not 16-byte aligned", what);
    LEA_CHECK((int64_t)v->D * v->H <= 65535, "%s: D*H = %lld exceeds the launch grid limit", what,
              (long long)v->D * v->H);
    return 0;
}
static int lea_check_slice(const lea_vol* v, int c0, int c, const char* what) {
    LEA_CHECK((c0 & 7) == 0 && (c & 7) == 0 && c > 0, "%s: channel slice [%d,+%d) must be 8-aligned", what, c0, c);
    LEA_CHECK(c0 >= 0 && c0 + c <= v->C, "%s: channel slice [%d,+%d) outside %d channels", what, c0, c, v->C);
    return 0;
}
static bool lea_same_space(const lea_vol* a, const lea_vol* b) {
    return a->B == b->B && a->D == b->D && a->H == b->H && a->W == b->W;
}

extern "C" int lea_cost_volume_f32(const float* x, const float* y, float* cost,
                                   int32_t B, int32_t C, int32_t H, int32_t W, int32_t D3, void* stream) {
    LEA_CHECK(B > 0 && C > 0 && H > 0 && W > 0 && D3 >= 0, "cost_volume_f32: bad shape");
    if (D3 == 0) return 0;                                   // int(maxdisp/3) == 0 -> empty volume, nothing to write
    LEA_CHECK(x && y && cost, "cost_volume_f32: null pointer");
    LEA_CHECK((int64_t)2 * C * D3 <= 65535 && B <= 65535, "cost_volume_f32: grid too large");
    const bool vec = (W % 4 == 0) && ((((uintptr_t)cost) & 15) == 0);
    if (vec) {
        const int hw4 = H * (W / 4);
        const int chunks = (hw4 + 1023) / 1024;
        LEA_LAUNCH(lea_cost_volume_f32_kernel, dim3(chunks, 2 * C * D3, B), dim3(256), 0, stream,
                   x, y, cost, C, H, W, D3, hw4, chunks);
    } else {
        LEA_LAUNCH(lea_cost_volume_f32_scalar_kernel, dim3((H * W + 255) / 256, 2 * C * D3, B), dim3(256), 0, stream,
                   x, y, cost, C, H, W, D3);
    }
    return LEA_POST_LAUNCH();
}

extern "C" int lea_cost_volume_planes(const float* x, const float* y, const lea_vol* vol, int32_t C, void* stream) {
    LEA_CHECK(x && y, "cost_volume_planes: null pointer");
    if (lea_check_vol(vol, "cost_volume_planes")) return 1;
    LEA_CHECK(C > 0 && (C & 7) == 0 && vol->C == 2 * C, "cost_volume_planes: volume must have 2C channels, C%%8==0");
    LEA_CHECK((int64_t)vol->B * (vol->C >> 3) <= 65535, "cost_volume_planes: grid too large");
    LEA_LAUNCH(lea_cost_volume_planes_kernel, dim3((vol->W + 255) / 256, vol->D * vol->H, vol->B * (vol->C >> 3)),
               dim3(256), 0, stream, x, y, *vol, C);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_pack_planes(const float* src, const lea_vol* dst, int32_t dst_c0, int32_t c, void* stream) {
    LEA_CHECK(src, "pack_planes: null pointer");
    if (lea_check_vol(dst, "pack_planes") || lea_check_slice(dst, dst_c0, c, "pack_planes")) return 1;
    LEA_CHECK((int64_t)dst->B * (c >> 3) <= 65535, "pack_planes: grid too large");
    LEA_LAUNCH(lea_pack_planes_kernel, dim3((dst->W + 255) / 256, dst->D * dst->H, dst->B * (c >> 3)), dim3(256), 0,
               stream, src, *dst, dst_c0, c);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_unpack_planes(const lea_vol* src, int32_t src_c0, int32_t c, float* dst, void* stream) {
    LEA_CHECK(dst, "unpack_planes: null pointer");
    if (lea_check_vol(src, "unpack_planes") || lea_check_slice(src, src_c0, c, "unpack_planes")) return 1;
    LEA_CHECK((int64_t)src->B * (c >> 3) <= 65535, "unpack_planes: grid too large");
    LEA_LAUNCH(lea_unpack_planes_kernel, dim3((src->W + 255) / 256, src->D * src->H, src->B * (c >> 3)), dim3(256), 0,
               stream, *src, src_c0, c, dst);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_trilinear_ac(const lea_vol* src, int32_t src_c0, const lea_vol* dst, int32_t dst_c0, int32_t c,
                                const float* bn_scale, const float* bn_shift, int32_t relu, void* stream) {
    LEA_CHECK((bn_scale == nullptr) == (bn_shift == nullptr), "trilinear_ac: bn_scale/bn_shift must come together");
    if (lea_check_vol(src, "trilinear_ac src") || lea_check_vol(dst, "trilinear_ac dst")) return 1;
    if (lea_check_slice(src, src_c0, c, "trilinear_ac src") || lea_check_slice(dst, dst_c0, c, "trilinear_ac dst"))
        return 1;
    LEA_CHECK(src->B == dst->B, "trilinear_ac: batch mismatch");
    LEA_CHECK((int64_t)dst->B * (c >> 3) <= 65535, "trilinear_ac: grid too large");
    const int nchunk = (dst->D + LEA_UP_DCH - 1) / LEA_UP_DCH;
    if (dst->D > src->D && dst->D >= 4 && (int64_t)nchunk * dst->H <= 65535) {
        // depth is up-sampled: depth-marching kernel (fewer loads per output)
        const dim3 grid((dst->W + 127) / 128, nchunk * dst->H, dst->B * (c >> 3));
        if (src->P == 2 && dst->P == 2)
            LEA_LAUNCH(lea_trilinear_ac_up_kernel<2>, grid, dim3(128), 0, stream, *src, src_c0, *dst, dst_c0, c, bn_scale, bn_shift, relu);
        else if (src->P == 3 && dst->P == 3)
            LEA_LAUNCH(lea_trilinear_ac_up_kernel<3>, grid, dim3(128), 0, stream, *src, src_c0, *dst, dst_c0, c, bn_scale, bn_shift, relu);
        else
            LEA_LAUNCH(lea_trilinear_ac_up_kernel<0>, grid, dim3(128), 0, stream, *src, src_c0, *dst, dst_c0, c, bn_scale, bn_shift, relu);
    } else {
        LEA_LAUNCH(lea_trilinear_ac_kernel, dim3((dst->W + 127) / 128, dst->D * dst->H, dst->B * (c >> 3)), dim3(128),
                   0, stream, *src, src_c0, *dst, dst_c0, c, bn_scale, bn_shift, relu);
    }
    return LEA_POST_LAUNCH();
}

template <int NT>
static int lea_launch_resample_conv1(const lea_vol* src, int32_t src_c0, int32_t c_in, const lea_rc_out* o, int32_t n_out,
                                     void* stream) {
    const size_t smem = (size_t)c_in * NT * sizeof(float);
#ifndef LEA_CPU_EMU
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(lea_resample_conv1_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    const lea_vol& dst = o[0].dst;
    LEA_LAUNCH(lea_resample_conv1_kernel<NT>, dim3((dst.W + 127) / 128, dst.D * dst.H, dst.B), dim3(128), smem, stream,
               *src, src_c0, c_in, o[0], n_out > 1 ? o[1] : o[0], n_out);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_resample_conv1x1(const lea_vol* src, int32_t src_c0, int32_t c_in, const lea_rc_out* outs, int32_t n_out,
                                    void* stream) {
    if (lea_check_vol(src, "resample_conv1x1 src") || lea_check_slice(src, src_c0, c_in, "resample_conv1x1 src")) return 1;
    LEA_CHECK(outs != nullptr && (n_out == 1 || n_out == 2), "resample_conv1x1: 1 or 2 consumers");
    int total = 0;
    for (int k = 0; k < n_out; ++k) {
        const lea_rc_out& o = outs[k];
        if (lea_check_vol(&o.dst, "resample_conv1x1 dst") || lea_check_slice(&o.dst, o.dst_c0, o.c_out, "resample_conv1x1 dst"))
            return 1;
        LEA_CHECK(o.weight != nullptr && (o.bn_scale == nullptr) == (o.bn_shift == nullptr), "resample_conv1x1: bad consumer");
        LEA_CHECK(o.dst.B == src->B && lea_same_space(&o.dst, &outs[0].dst), "resample_conv1x1: consumers must share the target geometry");
        total += o.c_out;
    }
    LEA_CHECK(total <= 64 && c_in <= 256, "resample_conv1x1: at most 64 output channels in total and 256 input channels");
    LEA_CHECK(outs[0].dst.B <= 65535, "resample_conv1x1: grid too large");
    if (total <= 16) return lea_launch_resample_conv1<16>(src, src_c0, c_in, outs, n_out, stream);
    if (total <= 32) return lea_launch_resample_conv1<32>(src, src_c0, c_in, outs, n_out, stream);
    return lea_launch_resample_conv1<64>(src, src_c0, c_in, outs, n_out, stream);
}

static int lea_check_conv(const lea_conv* p, const char* what) {
    LEA_CHECK(p != nullptr, "%s: null params", what);
    if (lea_check_vol(&p->src, what) || lea_check_slice(&p->src, p->src_c0, p->c_in, what)) return 1;
    LEA_CHECK(p->ksize == 1 || p->ksize == 3, "%s: kernel size %d not supported (1 or 3)", what, p->ksize);
    LEA_CHECK(p->c_out >= 1 && p->c_out <= 64, "%s: c_out %d outside 1..64", what, p->c_out);
    LEA_CHECK((p->bn_scale == nullptr) == (p->bn_shift == nullptr), "%s: bn_scale/bn_shift must come together", what);
    if (p->dst_f32 == nullptr) {
        if (lea_check_vol(&p->dst, what) || lea_check_slice(&p->dst, p->dst_c0, p->c_out, what)) return 1;
        LEA_CHECK(lea_same_space(&p->src, &p->dst), "%s: src/dst spatial shapes differ", what);
        if (p->has_res) {
            if (lea_check_vol(&p->res, what) || lea_check_slice(&p->res, p->res_c0, p->c_out, what)) return 1;
            LEA_CHECK(lea_same_space(&p->src, &p->res), "%s: src/res spatial shapes differ", what);
        }
    } else {
        LEA_CHECK(!p->has_res, "%s: residual add is not available with fp32 output", what);
    }
    return 0;
}

template <int NPAD>
static int lea_launch_conv_simt(const lea_conv* p, const float* weight, void* stream) {
    const int tiles = ((p->src.W + LEA_TW - 1) / LEA_TW) * ((p->src.H + LEA_TH - 1) / LEA_TH);
    const dim3 grid(tiles, p->src.D, p->src.B);
    if (p->ksize == 3) {
        const size_t smem = (size_t)(8 * 3 * (LEA_TH + 2) * (LEA_TW + 2) + 216 * NPAD) * sizeof(float);
#ifndef LEA_CPU_EMU
        if (smem > 48 * 1024)   // per device, cheap; DataParallel replicas live on several devices
            cudaFuncSetAttribute(lea_conv3_simt_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
        LEA_LAUNCH(lea_conv3_simt_kernel<NPAD>, grid, dim3(128), smem, stream, *p, weight);
    } else {
        const size_t smem = (size_t)p->c_in * NPAD * sizeof(float);
        LEA_CHECK(smem <= 48 * 1024, "conv3d_simt: 1x1x1 weights (%d x %d) do not fit shared memory", p->c_in, NPAD);
        LEA_LAUNCH(lea_conv1_simt_kernel<NPAD>, grid, dim3(128), smem, stream, *p, weight);
    }
    return LEA_POST_LAUNCH();
}

extern "C" int lea_conv3d_simt(const lea_conv* p, const float* weight, void* stream) {
    if (lea_check_conv(p, "conv3d_simt")) return 1;
    LEA_CHECK(weight != nullptr, "conv3d_simt: null weight");
    LEA_CHECK(p->src.B <= 65535 && p->src.D <= 65535, "conv3d_simt: grid too large");
    if (p->c_out <= 8)  return lea_launch_conv_simt<8>(p, weight, stream);
    if (p->c_out <= 16) return lea_launch_conv_simt<16>(p, weight, stream);
    if (p->c_out <= 32) return lea_launch_conv_simt<32>(p, weight, stream);
    return lea_launch_conv_simt<64>(p, weight, stream);
}

extern "C" int lea_disp_head(const float* mat, float* disp, int32_t B, int32_t D3, int32_t H3, int32_t W3,
                             int32_t maxdisp, void* stream) {
    LEA_CHECK(mat && disp, "disp_head: null pointer");
    LEA_CHECK(B > 0 && D3 > 0 && H3 > 0 && W3 > 0 && maxdisp > 0 && B <= 65535, "disp_head: bad shape");
    LEA_CHECK(H3 <= 65535, "disp_head: grid too large");
    // pass 1's blended logits stay in shared memory for pass 2 when a lane's quarter of the disparity range fits
    const int kchunk = (D3 + LEA_DH_PARTS - 1) / LEA_DH_PARTS;
    size_t smem = (size_t)kchunk * 9 * LEA_DH_CELLS * LEA_DH_PARTS * sizeof(float);
    int cache = (smem <= 200 * 1024) ? 1 : 0;
    if (!cache) smem = 0;
#ifndef LEA_CPU_EMU
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(lea_disp_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    LEA_LAUNCH(lea_disp_head_kernel, dim3((W3 + LEA_DH_CELLS - 1) / LEA_DH_CELLS, H3, B),
               dim3(LEA_DH_CELLS * LEA_DH_PARTS), smem, stream, mat, disp, D3, H3, W3, maxdisp, cache);
    return LEA_POST_LAUNCH();
}

extern "C" int64_t lea_head_taps_workspace_bytes(int32_t B, int32_t D1, int32_t H1, int32_t D, int32_t H, int32_t W) {
    if (B <= 0 || D1 <= 0 || H1 <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    (void)D;
    return (int64_t)sizeof(float) * B * D1 * W * ((int64_t)9 * H1 + (int64_t)3 * H);
}

extern "C" int lea_head_taps(const lea_vol* q, int32_t q_c0, float* mat, int32_t D, int32_t H, int32_t W,
                             float* workspace, void* stream) {
    LEA_CHECK(mat && workspace, "head_taps: null pointer");
    if (lea_check_vol(q, "head_taps")) return 1;
    LEA_CHECK((q_c0 & 7) == 0 && q_c0 >= 0 && q_c0 + 32 <= q->C, "head_taps: needs 27 tap channels padded to 32 at an "
              "8-aligned offset (slice [%d,+32) of %d channels)", q_c0, q->C);
    LEA_CHECK((D >= 2 * q->D - 1 || D == q->D) && (H >= 2 * q->H - 1 || H == q->H) && (W >= 2 * q->W - 1 || W == q->W),
              "head_taps: every axis must be up-sampled by >= 2x-1 (or kept): (%d,%d,%d) -> (%d,%d,%d)",
              q->D, q->H, q->W, D, H, W);
    LEA_CHECK((int64_t)q->D * H <= 65535 && (int64_t)D * H <= 65535 && q->B <= 65535, "head_taps: grid too large");
    float* R = workspace;
    float* S = workspace + (int64_t)q->B * 9 * q->D * q->H * W;
    const int gx = (W + 127) / 128;
    LEA_LAUNCH(lea_head_taps_w_kernel, dim3(gx, q->D * q->H, q->B), dim3(128), 0, stream, *q, q_c0, R, W);
    const bool vec = (W % 4 == 0) && ((((uintptr_t)workspace) & 15) == 0) && ((((uintptr_t)mat) & 15) == 0);
    if (vec) {
        const int gx4 = (W / 4 + 127) / 128;
        LEA_LAUNCH(lea_head_taps_h_kernel<4>, dim3(gx4, q->D * H, q->B), dim3(128), 0, stream, R, S, q->D, q->H, H, W);
        LEA_LAUNCH(lea_head_taps_d_kernel<4>, dim3(gx4, D * H, q->B), dim3(128), 0, stream, S, mat, q->D, D, H, W);
    } else {
        LEA_LAUNCH(lea_head_taps_h_kernel<1>, dim3(gx, q->D * H, q->B), dim3(128), 0, stream, R, S, q->D, q->H, H, W);
        LEA_LAUNCH(lea_head_taps_d_kernel<1>, dim3(gx, D * H, q->B), dim3(128), 0, stream, S, mat, q->D, D, H, W);
    }
    return LEA_POST_LAUNCH();
}

extern "C" int lea_stem0_assemble(const lea_vol* lmap, const lea_vol* abmap, const lea_vol* dst, int32_t dst_c0,
                                  int32_t c_out, const float* bn_scale, const float* bn_shift, int32_t relu,
                                  void* stream) {
    if (lea_check_vol(lmap, "stem0_assemble lmap") || lea_check_vol(abmap, "stem0_assemble abmap") ||
        lea_check_vol(dst, "stem0_assemble dst") || lea_check_slice(dst, dst_c0, c_out, "stem0_assemble dst"))
        return 1;
    LEA_CHECK((bn_scale == nullptr) == (bn_shift == nullptr), "stem0_assemble: bn_scale/bn_shift must come together");
    LEA_CHECK(lmap->D == 1 && abmap->D == 1 && lmap->C == c_out && abmap->C == 2 * c_out && lmap->B == dst->B &&
              abmap->B == dst->B && lmap->H == dst->H && abmap->H == dst->H && lmap->W == dst->W && abmap->W == dst->W,
              "stem0_assemble: maps must be (B, c_out | 2*c_out, 1, H, W) volumes matching dst");
    LEA_CHECK((int64_t)dst->B * (c_out >> 3) <= 65535, "stem0_assemble: grid too large");
    const int nchunk = (dst->D + LEA_AS_DCH - 1) / LEA_AS_DCH;
    LEA_LAUNCH(lea_stem0_assemble_kernel, dim3((dst->W + 127) / 128, nchunk * dst->H, dst->B * (c_out >> 3)), dim3(128),
               0, stream, *lmap, *abmap, *dst, dst_c0, c_out, bn_scale, bn_shift, relu);
    return LEA_POST_LAUNCH();
}

extern "C" int lea_disparity_regression(const float* p, float* out, int32_t B, int32_t maxdisp, int32_t H, int32_t W,
                                        void* stream) {
    LEA_CHECK(p && out, "disparity_regression: null pointer");
    LEA_CHECK(B > 0 && maxdisp > 0 && H > 0 && W > 0 && B <= 65535, "disparity_regression: bad shape");
    LEA_LAUNCH(lea_disparity_regression_kernel, dim3((H * W + 255) / 256, B), dim3(256), 0, stream,
               p, out, maxdisp, H * W);
    return LEA_POST_LAUNCH();
}


extern "C" int lea_feature_stem(const float* img, int32_t B, int32_t H, int32_t W,
                                const float* w0, const float* scale0, const float* shift0, int32_t c_mid,
                                const float* w1, const float* scale1, const float* shift1, int32_t c_out,
                                const lea_vol* dst, int32_t dst_c0, void* stream) {
    LEA_CHECK(img && w0 && scale0 && shift0 && w1 && scale1 && shift1, "feature_stem: null pointer");
    if (lea_check_vol(dst, "feature_stem") || lea_check_slice(dst, dst_c0, c_out, "feature_stem")) return 1;
    LEA_CHECK(c_mid >= 1 && c_mid <= LEA_FS_MAXMID && c_out <= LEA_FS_MAXOUT, "feature_stem: supports c_mid <= %d, c_out <= %d",
              LEA_FS_MAXMID, LEA_FS_MAXOUT);
    LEA_CHECK(dst->B == B && dst->D == 1 && dst->H == (H - 1) / 3 + 1 && dst->W == (W - 1) / 3 + 1,
              "feature_stem: output volume must be (B, c, 1, ceil(H/3), ceil(W/3))");
    LEA_CHECK(dst->H <= 65535 && B <= 65535, "feature_stem: grid too large");
    LEA_LAUNCH(lea_feature_stem_kernel, dim3((dst->W + 255) / 256, dst->H, B), dim3(128), 0, stream,
               img, H, W, w0, scale0, shift0, c_mid, w1, scale1, shift1, c_out, *dst, dst_c0);
    return LEA_POST_LAUNCH();
}
