/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product (sayuri_b200/) never links it.
 *
 * Plain-C restatement of the reference's CPU forward for ONE position (batch 1, native board size),
 * i.e. the algorithm of
 *     BlasForwardPipe::Forward               /root/reference/src/neural/blas/blas_forward_pipe.cc:314-563
 * together with the pieces it calls.  Every function cites the reference lines it follows.
 * Parity pin: tests/test_oracle.py checks this file against (a) golden vectors produced by the
 * UNMODIFIED reference compiled into oracle/_ref (tests/golden/make_golden.py) and (b) the
 * reference's independent PyTorch forward (train/torch/network.py:1121-1215) on the same weights.
 * The reference itself ships no tests / golden vectors, so those are the pins.
 *
 * fp32 arithmetic throughout, like the reference; summation order differs from Eigen's GEMM
 * (and the reference is built with -ffast-math), so agreement is ~1e-6 relative, not bit-exact.
 * Index work (canvas placement / crop, policy-channel select, symmetry tables) IS bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_INPUT_CHANNELS 43 /* network_basic.h:10 */
#define ORACLE_MAX_BLOCKS 128

/* activation.h:8-17 */
enum { ACT_IDENTITY = 0, ACT_RELU = 1, ACT_ELU = 2, ACT_SELU = 3, ACT_GELU = 4, ACT_MISH = 5, ACT_SWISH = 6, ACT_HARDSWISH = 7 };

typedef struct {
    int in, out, k;
    float* w; /* [out][in][k][k]  (OIHW, convolution.h:57-60) */
    float* b; /* [out] */
} conv_t;

typedef struct {
    int in, out;
    float* w; /* [out][in] row-major (blas.cc:155-162) */
    float* b;
} fc_t;

/* description.h:88-132 (BlockBasic) */
enum { BLK_RESIDUAL = 0, BLK_BOTTLENECK = 1, BLK_NESTED = 2, BLK_MIXER = 3 };

typedef struct {
    int type, inner; /* inner = bottleneck_channels / feedforward_channels (0 for a plain residual block) */
    conv_t conv1, conv2, conv3, conv4, pre, post;
    conv_t dw;       /* Mixer: depthwise k x k, w = [C][k][k] (in == 1 marks depthwise) */
    int apply_se, se_size;
    fc_t squeeze, excite;
} block_t;

typedef struct oracle_net {
    int version, input_channels, blocks, channels, P, V, act;
    conv_t input_conv;
    block_t* tower;
    conv_t p_hd_conv;
    int replk;               /* PolicyHeadType RepLK: depthwise k x k + 1x1 after the head-entry conv */
    conv_t p_dw_conv, p_pt_conv;
    fc_t p_inter_fc;
    conv_t prob_conv;
    fc_t pass_fc;
    conv_t v_hd_conv;
    fc_t v_inter_fc;
    conv_t v_ownership;
    fc_t v_misc;
} oracle_net;

/* ------------------------------------------------------------------------------------------ */
/* Weight-file reader: loader.cc:67-121 (Parse), :149-188 (ParseStruct), :190-239 (CheckMisc),  */
/* :628-773 (FillWeights), :833-898 (GetWeightsFromBuffer), parse_float.cc:5-33.                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const unsigned char* p;
    const unsigned char* end;
    int binary;
} rd_t;

static int rd_line(rd_t* r, char* buf, int cap) {
    if (r->p >= r->end) return 0;
    int n = 0;
    while (r->p < r->end && *r->p != '\n') {
        if (n < cap - 1) buf[n++] = (char)*r->p;
        r->p++;
    }
    if (r->p < r->end) r->p++;
    buf[n] = 0;
    return 1;
}

/* One tensor: text = one line of decimals parsed via double (loader.cc:849-896);
 * bin = little-endian float32 run terminated by the word 0xFFFFFFFF (loader.cc:836-847). */
static float* rd_tensor(rd_t* r, int expect, char* err, int errlen) {
    float* t = (float*)malloc(sizeof(float) * (size_t)(expect > 0 ? expect : 1));
    int n = 0;
    if (r->binary) {
        for (;;) {
            if (r->p + 4 > r->end) { snprintf(err, errlen, "truncated binary tensor"); free(t); return NULL; }
            uint32_t u;
            memcpy(&u, r->p, 4);
            r->p += 4;
            if (u == 0xffffffffu) break;
            if (n < expect) memcpy(&t[n], &u, 4);
            n++;
        }
    } else {
        const char* s = (const char*)r->p;
        const char* e = s;
        while ((const unsigned char*)e < r->end && *e != '\n') e++;
        size_t len = (size_t)(e - s);
        char* line = (char*)malloc(len + 1);
        memcpy(line, s, len);
        line[len] = 0;
        char* q = line;
        for (;;) {
            char* endp;
            double v = strtod(q, &endp);
            if (endp == q) break;
            if (n < expect) t[n] = (float)v;
            n++;
            q = endp;
        }
        free(line);
        r->p = (const unsigned char*)e;
        if (r->p < r->end) r->p++;
    }
    if (n != expect) {
        snprintf(err, errlen, "tensor size mismatch: expect %d got %d", expect, n);
        free(t);
        return NULL;
    }
    return t;
}

typedef struct { char kind; int d[3]; } shape_t; /* 'C' conv, 'D' depthwise, 'B' batchnorm, 'F' fc */

/* BN "std" tensor -> 1/std (v>=2) or 1/sqrt(var+1e-5) (v1): description.cc:70-85, description.h:44-54 */
static void bn_to_scale(float* s, int n, int v1) {
    for (int i = 0; i < n; ++i) s[i] = v1 ? 1.0f / sqrtf(s[i] + 1e-5f) : 1.0f / s[i];
}

/* Fold BN into the conv: loader.cc:776-789.  bias = (bias - mean) * scale ; W[o,...] *= scale[o]. */
static void fold_bn(conv_t* c, const float* mean, const float* scale) {
    size_t stride = (size_t)c->in * c->k * c->k;
    for (int o = 0; o < c->out; ++o) {
        c->b[o] -= mean[o];
        for (size_t k = 0; k < stride; ++k) c->w[stride * o + k] *= scale[o];
        c->b[o] *= scale[o];
    }
}

static int rd_conv(rd_t* r, const shape_t* sh, conv_t* c, char* err, int errlen) {
    if (sh->kind != 'C' && sh->kind != 'D') { snprintf(err, errlen, "expected Convolution layer"); return -1; }
    c->in = sh->d[0]; c->out = sh->d[1]; c->k = sh->d[2];
    if (sh->kind == 'D') {   /* "DepthwiseConvolution 1 C k" (network.py writer): one k x k filter per channel (convolution.cc:27-62) */
        if (c->in != 1 && c->in != c->out) { snprintf(err, errlen, "depthwise convolution shape is wrong"); return -1; }
        c->in = 1;
    }
    c->w = rd_tensor(r, c->in * c->out * c->k * c->k, err, errlen);
    if (!c->w) return -1;
    c->b = rd_tensor(r, c->out, err, errlen);
    return c->b ? 0 : -1;
}

static int rd_conv_bn(rd_t* r, const shape_t* sh, conv_t* c, int v1, char* err, int errlen) {
    if (rd_conv(r, &sh[0], c, err, errlen)) return -1;
    if (sh[1].kind != 'B' || sh[1].d[0] != c->out) { snprintf(err, errlen, "expected BatchNorm %d", c->out); return -1; }
    float* mean = rd_tensor(r, c->out, err, errlen);
    if (!mean) return -1;
    float* sd = rd_tensor(r, c->out, err, errlen);
    if (!sd) { free(mean); return -1; }
    bn_to_scale(sd, c->out, v1);
    fold_bn(c, mean, sd);
    free(mean);
    free(sd);
    return 0;
}

static int rd_fc(rd_t* r, const shape_t* sh, fc_t* f, char* err, int errlen) {
    if (sh->kind != 'F') { snprintf(err, errlen, "expected FullyConnect layer"); return -1; }
    f->in = sh->d[0]; f->out = sh->d[1];
    f->w = rd_tensor(r, f->in * f->out, err, errlen);
    if (!f->w) return -1;
    f->b = rd_tensor(r, f->out, err, errlen);
    return f->b ? 0 : -1;
}

static int act_from_name(const char* s) { /* activation.h:19-41 */
    char b[32]; int i = 0;
    for (; s[i] && i < 31; ++i) b[i] = (char)((s[i] >= 'A' && s[i] <= 'Z') ? s[i] + 32 : s[i]);
    b[i] = 0;
    const char* names[] = {"identity", "relu", "elu", "selu", "gelu", "mish", "swish", "hardswish"};
    for (int k = 0; k < 8; ++k) if (!strcmp(b, names[k])) return k;
    return -1;
}

void oracle_free(oracle_net* n);

oracle_net* oracle_load(const char* path, char* err, int errlen) {
    FILE* f = fopen(path, "rb");
    if (!f) { snprintf(err, errlen, "cannot open %s", path); return NULL; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char* data = (unsigned char*)malloc((size_t)sz + 1);
    if (fread(data, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(data); snprintf(err, errlen, "short read"); return NULL; }
    fclose(f);
    rd_t r = {data, data + sz, 0};
    char line[512];
    oracle_net* n = (oracle_net*)calloc(1, sizeof(oracle_net));
    n->version = 1;          /* loader.cc:198 */
    n->act = ACT_RELU;       /* loader.cc:261-265 */
    int residual_blocks = -1, channels = -1, P = -1, V = -1, in_ch = -1;
    char stack[ORACLE_MAX_BLOCKS][64];
    int n_stack = 0;
    shape_t* shapes = (shape_t*)calloc(4096, sizeof(shape_t));
    int n_shapes = 0;
    int ok = 0;

    if (!rd_line(&r, line, sizeof line) || strncmp(line, "get main", 8)) { snprintf(err, errlen, "weights file format is not acceptable"); goto fail; }
    while (rd_line(&r, line, sizeof line)) {
        char a[64] = {0}, b[64] = {0};
        sscanf(line, "%63s %63s", a, b);
        if (strcmp(a, "get")) continue;
        if (!strcmp(b, "info")) {
            while (rd_line(&r, line, sizeof line)) {
                char k[64] = {0}, v[128] = {0};
                if (sscanf(line, "%63s %127s", k, v) < 1) continue;
                if (k[0] == '#') continue;
                if (!strcmp(k, "end")) break;
                if (!strcmp(k, "Version")) n->version = atoi(v);
                else if (!strcmp(k, "FloatType")) r.binary = !strcmp(v, "float32bin");
                else if (!strcmp(k, "InputChannels")) in_ch = atoi(v);
                else if (!strcmp(k, "ResidualChannels")) channels = atoi(v);
                else if (!strcmp(k, "ResidualBlocks")) residual_blocks = atoi(v);
                else if (!strcmp(k, "PolicyHeadChannels") || !strcmp(k, "PolicyExtract")) P = atoi(v);
                else if (!strcmp(k, "ValueHeadChannels") || !strcmp(k, "ValueExtract")) V = atoi(v);
                else if (!strcmp(k, "PolicyHeadType")) {   /* loader.cc:245-259 */
                    if (!strcasecmp(v, "replk")) n->replk = 1;
                    else if (strcasecmp(v, "normal")) { snprintf(err, errlen, "policy head type %s is outside the oracle's scope", v); goto fail; }
                } else if (!strcmp(k, "ActivationFunction")) {
                    n->act = act_from_name(v);
                    if (n->act < 0) { snprintf(err, errlen, "Unknown activation type."); goto fail; }
                }
            }
        } else if (!strcmp(b, "stack")) {
            while (rd_line(&r, line, sizeof line)) {
                char k[64] = {0};
                if (sscanf(line, "%63s", k) < 1) continue;
                if (k[0] == '#') continue;
                if (!strcmp(k, "end")) break;
                if (n_stack < ORACLE_MAX_BLOCKS) strcpy(stack[n_stack++], k);
            }
        } else if (!strcmp(b, "struct")) {
            while (rd_line(&r, line, sizeof line)) {
                char k[64] = {0};
                int d0 = 0, d1 = 0, d2 = 0;
                int c = sscanf(line, "%63s %d %d %d", k, &d0, &d1, &d2);
                if (c < 1 || k[0] == '#') continue;
                if (!strcmp(k, "end")) break;
                shape_t* s = &shapes[n_shapes++];
                s->d[0] = d0; s->d[1] = d1; s->d[2] = d2;
                if (!strcmp(k, "Convolution") && c == 4) s->kind = 'C';
                else if (!strcmp(k, "DepthwiseConvolution") && c == 4) s->kind = 'D';
                else if (!strcmp(k, "BatchNorm") && c == 2) s->kind = 'B';
                else if (!strcmp(k, "FullyConnect") && c == 3) s->kind = 'F';
                else { snprintf(err, errlen, "layer shape is error"); goto fail; }
            }
        } else if (!strcmp(b, "parameters")) {
            break;
        }
    }
    if (n->version >= 6) { snprintf(err, errlen, "do not support this version"); goto fail; }
    if (n->version < 3) { snprintf(err, errlen, "v1/v2 nets (38 planes) are outside the oracle's scope"); goto fail; }
    n->input_channels = ORACLE_INPUT_CHANNELS;
    if (in_ch != n->input_channels) { snprintf(err, errlen, "the number of input channels is wrong"); goto fail; }
    if (residual_blocks < 0 || channels <= 0 || P <= 0 || V <= 0) { snprintf(err, errlen, "missing info fields"); goto fail; }
    n->blocks = residual_blocks; n->channels = channels; n->P = P; n->V = V;
    if (n_stack == 0) { /* loader.cc:270-292: infer ResidualBlock[-SE] from the struct list */
        int inner = 0;
        for (int b = 0; b < residual_blocks; ++b) {
            strcpy(stack[b], "ResidualBlock");
            inner += 4;
            if (shapes[inner + 2].kind == 'F') { strcat(stack[b], "-SE"); inner += 2; }
        }
        n_stack = residual_blocks;
    }
    if (n_stack != residual_blocks) { snprintf(err, errlen, "stack size != ResidualBlocks"); goto fail; }
    {
        const int v1 = (n->version == 1);
        int off = 0;
        if (rd_conv_bn(&r, &shapes[off], &n->input_conv, v1, err, errlen)) goto fail;
        off += 2;
        if (n->input_conv.in != n->input_channels || n->input_conv.out != channels || n->input_conv.k != 3) { snprintf(err, errlen, "the input layers are wrong"); goto fail; }
        n->tower = (block_t*)calloc((size_t)(residual_blocks > 0 ? residual_blocks : 1), sizeof(block_t));
        for (int b = 0; b < residual_blocks; ++b) {
            block_t* blk = &n->tower[b];
            char* dash = strchr(stack[b], '-');
            int se = 0;
            if (dash) { if (strcmp(dash + 1, "SE")) { snprintf(err, errlen, "block component %s outside the oracle's scope", dash + 1); goto fail; } *dash = 0; se = 1; }
            if (!strcmp(stack[b], "ResidualBlock")) {
                /* loader.cc:385-415: conv1,bn1,conv2,bn2 [, squeeze fc, excite fc] */
                blk->type = BLK_RESIDUAL;
                if (rd_conv_bn(&r, &shapes[off], &blk->conv1, v1, err, errlen)) goto fail;
                off += 2;
                if (rd_conv_bn(&r, &shapes[off], &blk->conv2, v1, err, errlen)) goto fail;
                off += 2;
                if (blk->conv1.k != 3 || blk->conv2.k != 3 || blk->conv1.in != channels || blk->conv1.out != channels || blk->conv2.in != channels || blk->conv2.out != channels) { snprintf(err, errlen, "the Nth residual block is wrong"); goto fail; }
            } else if (!strcmp(stack[b], "BottleneckBlock") || !strcmp(stack[b], "NestedBottleneckBlock")) {
                /* loader.cc:416-465 (pre 1x1, conv1, conv2, post 1x1) and :466-555 (pre, conv1..conv4, post) */
                const int nested = stack[b][0] == 'N';
                blk->type = nested ? BLK_NESTED : BLK_BOTTLENECK;
                conv_t* seq[6];
                int ns = 0;
                seq[ns++] = &blk->pre; seq[ns++] = &blk->conv1; seq[ns++] = &blk->conv2;
                if (nested) { seq[ns++] = &blk->conv3; seq[ns++] = &blk->conv4; }
                seq[ns++] = &blk->post;
                for (int q = 0; q < ns; ++q) {
                    if (rd_conv_bn(&r, &shapes[off], seq[q], v1, err, errlen)) goto fail;
                    off += 2;
                }
                blk->inner = blk->pre.out;
                if (blk->pre.k != 1 || blk->post.k != 1 || blk->pre.in != channels || blk->post.out != channels || blk->post.in != blk->inner) { snprintf(err, errlen, "the outer channels of bottleneck block is wrong"); goto fail; }
                for (int q = 1; q < ns - 1; ++q)
                    if (seq[q]->k != 3 || seq[q]->in != blk->inner || seq[q]->out != blk->inner) { snprintf(err, errlen, "the inner channels of bottleneck block is wrong"); goto fail; }
            } else if (!strcmp(stack[b], "MixerBlock")) {
                /* loader.cc:556-607: depthwise k x k + bn, 1x1 C->F + bn, 1x1 F->C + bn */
                blk->type = BLK_MIXER;
                if (rd_conv_bn(&r, &shapes[off], &blk->dw, v1, err, errlen)) goto fail;
                off += 2;
                if (rd_conv_bn(&r, &shapes[off], &blk->conv1, v1, err, errlen)) goto fail;
                off += 2;
                if (rd_conv_bn(&r, &shapes[off], &blk->conv2, v1, err, errlen)) goto fail;
                off += 2;
                blk->inner = blk->conv1.out;
                if (blk->dw.in != 1 || blk->dw.out != channels || blk->conv1.k != 1 || blk->conv2.k != 1 || blk->conv1.in != channels || blk->conv2.in != blk->inner || blk->conv2.out != channels) { snprintf(err, errlen, "the channels of mixer block is wrong"); goto fail; }
            } else { snprintf(err, errlen, "block type %s is outside the oracle's scope", stack[b]); goto fail; }
            if (se) {
                if (rd_fc(&r, &shapes[off], &blk->squeeze, err, errlen)) goto fail;
                off += 1;
                if (rd_fc(&r, &shapes[off], &blk->excite, err, errlen)) goto fail;
                off += 1;
                blk->apply_se = 1;
                blk->se_size = blk->squeeze.out;
                if (blk->squeeze.in != 3 * channels || blk->excite.in != blk->se_size || blk->excite.out != 2 * channels) { snprintf(err, errlen, "the SE unit size is wrong"); goto fail; }
            }
        }
        /* policy head, loader.cc:684-729 */
        if (rd_conv_bn(&r, &shapes[off], &n->p_hd_conv, v1, err, errlen)) goto fail;
        off += 2;
        if (n->replk) {   /* loader.cc:691-702 */
            if (rd_conv_bn(&r, &shapes[off], &n->p_dw_conv, v1, err, errlen)) goto fail;
            off += 2;
            if (rd_conv_bn(&r, &shapes[off], &n->p_pt_conv, v1, err, errlen)) goto fail;
            off += 2;
            if (n->p_dw_conv.in != 1 || n->p_dw_conv.out != P || n->p_pt_conv.k != 1 || n->p_pt_conv.in != P || n->p_pt_conv.out != P) { snprintf(err, errlen, "the RepLK policy head is wrong"); goto fail; }
        }
        if (rd_fc(&r, &shapes[off++], &n->p_inter_fc, err, errlen)) goto fail;
        if (rd_conv(&r, &shapes[off++], &n->prob_conv, err, errlen)) goto fail;
        if (rd_fc(&r, &shapes[off++], &n->pass_fc, err, errlen)) goto fail;
        if (n->p_hd_conv.k != 1 || n->prob_conv.k != 1 || n->prob_conv.out != 5 || n->p_inter_fc.in != 3 * P || n->p_inter_fc.out != P || n->pass_fc.in != P || n->pass_fc.out != 5 || n->p_hd_conv.out != P) { snprintf(err, errlen, "the policy head is wrong"); goto fail; }
        /* value head, loader.cc:731-761 */
        if (rd_conv_bn(&r, &shapes[off], &n->v_hd_conv, v1, err, errlen)) goto fail;
        off += 2;
        if (rd_fc(&r, &shapes[off++], &n->v_inter_fc, err, errlen)) goto fail;
        if (rd_conv(&r, &shapes[off++], &n->v_ownership, err, errlen)) goto fail;
        if (rd_fc(&r, &shapes[off++], &n->v_misc, err, errlen)) goto fail;
        if (n->v_hd_conv.k != 1 || n->v_ownership.k != 1 || n->v_ownership.out != 1 || n->v_inter_fc.in != 3 * V || n->v_inter_fc.out != 3 * V || n->v_misc.in != 3 * V || n->v_misc.out != 15 || n->v_hd_conv.out != V) { snprintf(err, errlen, "the value head is wrong"); goto fail; }
        if (off != n_shapes) { snprintf(err, errlen, "struct has %d layers, consumed %d", n_shapes, off); goto fail; }
        /* loader.cc:763-768: the next word must be "end" */
        if (!rd_line(&r, line, sizeof line) || strncmp(line, "end", 3)) { snprintf(err, errlen, "weights file format is not acceptable"); goto fail; }
    }
    ok = 1;
fail:
    free(shapes);
    free(data);
    if (!ok) { oracle_free(n); return NULL; }
    return n;
}

static void free_conv(conv_t* c) { free(c->w); free(c->b); }
static void free_fc(fc_t* f) { free(f->w); free(f->b); }

void oracle_free(oracle_net* n) {
    if (!n) return;
    free_conv(&n->input_conv);
    if (n->tower) {
        for (int b = 0; b < n->blocks; ++b) {
            free_conv(&n->tower[b].conv1); free_conv(&n->tower[b].conv2); free_conv(&n->tower[b].conv3);
            free_conv(&n->tower[b].conv4); free_conv(&n->tower[b].pre); free_conv(&n->tower[b].post); free_conv(&n->tower[b].dw);
            free_fc(&n->tower[b].squeeze); free_fc(&n->tower[b].excite);
        }
        free(n->tower);
    }
    free_conv(&n->p_hd_conv); free_conv(&n->p_dw_conv); free_conv(&n->p_pt_conv); free_fc(&n->p_inter_fc); free_conv(&n->prob_conv); free_fc(&n->pass_fc);
    free_conv(&n->v_hd_conv); free_fc(&n->v_inter_fc); free_conv(&n->v_ownership); free_fc(&n->v_misc);
    free(n);
}

/* {version, input_channels, blocks, channels, P, V, activation, n_se_blocks} */
int oracle_info(const oracle_net* n, int* out8) {
    out8[0] = n->version; out8[1] = n->input_channels; out8[2] = n->blocks; out8[3] = n->channels;
    out8[4] = n->P; out8[5] = n->V; out8[6] = n->act;
    int nse = 0;
    for (int b = 0; b < n->blocks; ++b) nse += n->tower[b].apply_se;
    out8[7] = nse;
    return 0;
}

/* Folded tensors in loader order (for comparing the product's own loader with this one, bit-exact).
 * idx enumerates: input_conv, [(pre,) conv1, conv2, (conv3, conv4,) (post,) (squeeze, excite)] per block, p_hd, p_inter, prob, pass,
 * v_hd, v_inter, own, misc; which = 0 weights, 1 biases.  Returns element count, or -1 past the end. */
int oracle_get_tensor(const oracle_net* n, int idx, int which, const float** out) {
    int i = 0;
#define EMIT_CONV(c) do { if (i++ == idx) { *out = which ? (c).b : (c).w; return which ? (c).out : (c).out * (c).in * (c).k * (c).k; } } while (0)
#define EMIT_FC(f) do { if (i++ == idx) { *out = which ? (f).b : (f).w; return which ? (f).out : (f).out * (f).in; } } while (0)
    EMIT_CONV(n->input_conv);
    for (int b = 0; b < n->blocks; ++b) {
        const int ty = n->tower[b].type;
        if (ty == BLK_MIXER) EMIT_CONV(n->tower[b].dw);
        if (ty == BLK_BOTTLENECK || ty == BLK_NESTED) EMIT_CONV(n->tower[b].pre);
        EMIT_CONV(n->tower[b].conv1);
        EMIT_CONV(n->tower[b].conv2);
        if (ty == BLK_NESTED) { EMIT_CONV(n->tower[b].conv3); EMIT_CONV(n->tower[b].conv4); }
        if (ty == BLK_BOTTLENECK || ty == BLK_NESTED) EMIT_CONV(n->tower[b].post);
        if (n->tower[b].apply_se) { EMIT_FC(n->tower[b].squeeze); EMIT_FC(n->tower[b].excite); }
    }
    EMIT_CONV(n->p_hd_conv);
    if (n->replk) { EMIT_CONV(n->p_dw_conv); EMIT_CONV(n->p_pt_conv); }
    EMIT_FC(n->p_inter_fc); EMIT_CONV(n->prob_conv); EMIT_FC(n->pass_fc);
    EMIT_CONV(n->v_hd_conv); EMIT_FC(n->v_inter_fc); EMIT_CONV(n->v_ownership); EMIT_FC(n->v_misc);
#undef EMIT_CONV
#undef EMIT_FC
    return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* Ops                                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* activation.h:43-59 (expf/logf/tanhf in fp32, same formulas) */
static float activate(float x, int act) {
    switch (act) {
        case ACT_RELU: return x > 0.f ? x : 0.f;
        case ACT_ELU: return x > 0.f ? x : (expf(x) - 1);
        case ACT_SELU: return x > 0.f ? (1.05070098f * x) : (1.05070098f * 1.67326324f * (expf(x) - 1.0f));
        case ACT_GELU: return (float)(0.5f * x * (1.0f + tanhf((float)(0.7978845608028654f * (x + 0.044715 * x * x * x)))));
        case ACT_MISH: return x * tanhf(logf(1.0f + expf(x)));
        case ACT_SWISH: return x / (1.0f + expf(-x));
        case ACT_HARDSWISH: return x >= 3.f ? x : x <= -3.f ? 0.f : (x * (x + 3.0f) / 6.0f);
        default: return x;
    }
}

/* 3x3 / 1x1 same-pad cross-correlation, stride 1, no bias: the product the reference forms with
 * Im2col + ConvolutionSgemm (convolution.h:41-125, convolution.cc:3-25): out[o][p] = sum_{c,ky,kx}
 * W[o][c][ky][kx] * in[c][y+ky-pad][x+kx-pad], zero outside the board. */
static void conv_forward(const conv_t* c, int bs, const float* in, float* out) {
    const int s = bs * bs, pad = c->k / 2;
    memset(out, 0, sizeof(float) * (size_t)c->out * s);
    for (int o = 0; o < c->out; ++o) {
        float* op = out + (size_t)o * s;
        for (int ci = 0; ci < c->in; ++ci) {
            const float* ip = in + (size_t)ci * s;
            const float* wp = c->w + ((size_t)o * c->in + ci) * c->k * c->k;
            for (int ky = 0; ky < c->k; ++ky) {
                for (int kx = 0; kx < c->k; ++kx) {
                    const float w = wp[ky * c->k + kx];
                    const int dy = ky - pad, dx = kx - pad;
                    const int y0 = dy < 0 ? -dy : 0, y1 = dy > 0 ? bs - dy : bs;
                    const int x0 = dx < 0 ? -dx : 0, x1 = dx > 0 ? bs - dx : bs;
                    for (int y = y0; y < y1; ++y) {
                        float* orow = op + y * bs;
                        const float* irow = ip + (y + dy) * bs + dx;
                        for (int x = x0; x < x1; ++x) orow[x] += w * irow[x];
                    }
                }
            }
        }
    }
}

/* AddSpatialBiases::Forward, biases.cc:14-45: x + bias[c] (+ residual), then activation. */
static void add_spatial_biases(int bs, int channels, float* x, const float* bias, const float* residual, int act) {
    const int s = bs * bs;
    for (int c = 0; c < channels; ++c) {
        const float b = bias ? bias[c] : 0.0f;
        for (int i = 0; i < s; ++i) {
            float v = x[(size_t)c * s + i] + b;
            if (residual) v += residual[(size_t)c * s + i];
            x[(size_t)c * s + i] = activate(v, act);
        }
    }
}

/* DepthwiseConvolution::Forward, convolution.cc:27-62: per-channel k x k cross-correlation, zero outside the board. */
static void dwconv_forward(const conv_t* c, int bs, const float* in, float* out) {
    const int s = bs * bs, k = c->k, pad = k / 2;
    for (int ch = 0; ch < c->out; ++ch) {
        const float* ip = in + (size_t)ch * s;
        const float* wp = c->w + (size_t)ch * k * k;
        for (int y = 0; y < bs; ++y)
            for (int x = 0; x < bs; ++x) {
                float v = 0.f;
                for (int ky = 0; ky < k; ++ky)
                    for (int kx = 0; kx < k; ++kx) {
                        const int yy = y + ky - pad, xx = x + kx - pad;
                        if (yy >= 0 && yy < bs && xx >= 0 && xx < bs) v += ip[yy * bs + xx] * wp[ky * k + kx];
                    }
                out[(size_t)ch * s + y * bs + x] = v;
            }
    }
}

/* AddSpatialBiasesPost::Forward, biases.cc:47-77: activation FIRST, then the residual. */
static void add_spatial_biases_post(int bs, int channels, float* x, const float* bias, const float* residual, int act) {
    const int s = bs * bs;
    for (int c = 0; c < channels; ++c)
        for (int i = 0; i < s; ++i) {
            float v = activate(x[(size_t)c * s + i] + bias[c], act);
            if (residual) v += residual[(size_t)c * s + i];
            x[(size_t)c * s + i] = v;
        }
}

/* FullyConnect::Forward, fullyconnect.cc:7-19 + AddVectorBiases biases.cc:79-89 */
static void fc_forward(const fc_t* f, const float* in, float* out, int act) {
    for (int o = 0; o < f->out; ++o) {
        float acc = 0.f;
        for (int i = 0; i < f->in; ++i) acc += f->w[(size_t)o * f->in + i] * in[i];
        out[o] = activate(f->b[o] + acc, act);
    }
}

#define K_AVG_BSIZE 14.0f      /* se_unit.h:17-20: (19+9)/2 */
#define K_BSIZE_VARIANCE 0.1f

/* GlobalPooling<false>::Forward, se_unit.cc:9-37: [mean, mean*(N-14)/10, max(start -5000)] */
static void global_pool(int bs, int channels, const float* x, float* out) {
    const int s = bs * bs;
    const float b_coeff = ((float)bs - K_AVG_BSIZE) / 10.f;
    for (int c = 0; c < channels; ++c) {
        float sum = 0.f, mx = -5000.0f;
        for (int i = 0; i < s; ++i) {
            const float v = x[(size_t)c * s + i];
            sum += v;
            mx = v > mx ? v : mx;
        }
        const float mean = sum / (float)s;
        out[c] = mean;
        out[c + channels] = mean * b_coeff;
        out[c + 2 * channels] = mx;
    }
}

/* GlobalPooling<true>::Forward (value head), se_unit.cc:39-68 */
static void global_pool_value(int bs, int channels, const float* x, float* out) {
    const int s = bs * bs;
    const float b_diff = (float)bs - K_AVG_BSIZE;
    const float c0 = b_diff / 10.f;
    const float c1 = b_diff * b_diff / 100.f - K_BSIZE_VARIANCE;
    for (int c = 0; c < channels; ++c) {
        float sum = 0.f;
        for (int i = 0; i < s; ++i) sum += x[(size_t)c * s + i];
        const float mean = sum / (float)s;
        out[c] = mean;
        out[c + channels] = mean * c0;
        out[c + 2 * channels] = mean * c1;
    }
}

/* SEUnit::Forward + SEProcess, se_unit.cc:70-128 */
static void se_unit(int bs, int channels, const block_t* blk, float* x, const float* residual, int act) {
    const int s = bs * bs;
    float* pool = (float*)malloc(sizeof(float) * 3 * (size_t)channels);
    float* h = (float*)malloc(sizeof(float) * (size_t)blk->se_size);
    global_pool(bs, channels, x, pool);
    fc_forward(&blk->squeeze, pool, h, act);
    fc_forward(&blk->excite, h, pool, ACT_IDENTITY); /* 2C outputs reuse `pool` (3C) like the reference */
    for (int c = 0; c < channels; ++c) {
        const float gamma = 1.0f / (1.0f + expf(-pool[c]));
        const float beta = pool[c + channels];
        for (int i = 0; i < s; ++i) {
            float v = gamma * x[(size_t)c * s + i] + beta;
            if (residual) v += residual[(size_t)c * s + i];
            x[(size_t)c * s + i] = activate(v, act);
        }
    }
    free(pool);
    free(h);
}

/*
 * BlasForwardPipe::Forward (blas_forward_pipe.cc:314-563) + FillOutputs (:565-619, v3..v5 branch).
 * planes: 43*bs*bs floats NCHW at native board size (encoder.cc:31-50).
 * out (2*s + 8 floats): prob[s] (channel `offset` of the 5 policy planes, raw logits) | ownership[s] (raw)
 *                       | pass[offset], wdl0, wdl1, wdl2, stm, final_score(misc[8]), q_err(misc[13]), score_err(misc[14])
 * trunk (optional, C*s floats NCHW): tower output, for layer-level debugging of the CUDA path.
 */
int oracle_forward_trace(const oracle_net* n, const float* planes, int bs, int offset, float* out, float* trunk,
                         float* all_prob /* optional 5*s + 5 pass + 15 misc */) {
    if (offset < 0 || offset > 4 || bs < 1 || bs > 25) return -1;
    const int s = bs * bs, C = n->channels, act = n->act;
    float* x = (float*)malloc(sizeof(float) * (size_t)C * s);
    float* t = (float*)malloc(sizeof(float) * (size_t)C * s);
    float* u = (float*)malloc(sizeof(float) * (size_t)C * s);
    /* input layers :373-383 */
    conv_forward(&n->input_conv, bs, planes, x);
    add_spatial_biases(bs, C, x, n->input_conv.b, NULL, act);
    /* tower :386-424 */
    for (int b = 0; b < n->blocks; ++b) {
        const block_t* blk = &n->tower[b];
        const conv_t* last;     /* the conv whose output joins the skip connection (or feeds the SE unit) */
        if (blk->type == BLK_MIXER) {
            /* MixerBlockForward :265-312: y = act(dw(x) + b) + x ; out = conv2(act(conv1(y))) (+ y, act) */
            const int F = blk->inner;
            float* f = (float*)malloc(sizeof(float) * (size_t)F * s);
            dwconv_forward(&blk->dw, bs, x, t);
            add_spatial_biases_post(bs, C, t, blk->dw.b, x, act);
            conv_forward(&blk->conv1, bs, t, f);
            add_spatial_biases(bs, F, f, blk->conv1.b, NULL, act);
            conv_forward(&blk->conv2, bs, f, u);
            free(f);
            /* the skip of this block (and of its SE unit, :408-420) is y, not x */
            if (blk->apply_se) {
                add_spatial_biases(bs, C, u, blk->conv2.b, NULL, ACT_IDENTITY);
                se_unit(bs, C, blk, u, t, act);
            } else {
                add_spatial_biases(bs, C, u, blk->conv2.b, t, act);
            }
            float* tmp = x; x = u; u = tmp;
            continue;
        }
        if (blk->type == BLK_RESIDUAL) {
            /* ResidualBlockForward :46-88 */
            conv_forward(&blk->conv1, bs, x, t);
            add_spatial_biases(bs, C, t, blk->conv1.b, NULL, act);
            conv_forward(&blk->conv2, bs, t, u);
            last = &blk->conv2;
        } else {
            /* BottleneckBlockForward :90-162 / NestedBottleneckBlockForward :164-263 */
            const int I = blk->inner;
            float* a = (float*)malloc(sizeof(float) * (size_t)I * s);
            float* c1 = (float*)malloc(sizeof(float) * (size_t)I * s);
            float* c2 = (float*)malloc(sizeof(float) * (size_t)I * s);
            conv_forward(&blk->pre, bs, x, a);                          /* pre-bottleneck 1x1 */
            add_spatial_biases(bs, I, a, blk->pre.b, NULL, act);
            conv_forward(&blk->conv1, bs, a, c1);
            add_spatial_biases(bs, I, c1, blk->conv1.b, NULL, act);
            conv_forward(&blk->conv2, bs, c1, c2);
            if (blk->type == BLK_BOTTLENECK) {
                add_spatial_biases(bs, I, c2, blk->conv2.b, NULL, act);                 /* :128-129 no skip */
            } else {
                add_spatial_biases(bs, I, c2, blk->conv2.b, a, act);                    /* :218-219 + residual1 */
                conv_forward(&blk->conv3, bs, c2, c1);
                add_spatial_biases(bs, I, c1, blk->conv3.b, NULL, act);
                conv_forward(&blk->conv4, bs, c1, a);
                add_spatial_biases(bs, I, a, blk->conv4.b, c2, act);                    /* :246-247 + residual1 */
                float* sw = a; a = c2; c2 = sw;
            }
            conv_forward(&blk->post, bs, c2, u);                        /* post-bottleneck 1x1 */
            last = &blk->post;
            free(a); free(c1); free(c2);
        }
        if (blk->apply_se) {
            add_spatial_biases(bs, C, u, last->b, NULL, ACT_IDENTITY);
            se_unit(bs, C, blk, u, x, act);
        } else {
            add_spatial_biases(bs, C, u, last->b, x, act);
        }
        float* tmp = x; x = u; u = tmp;
    }
    if (trunk) memcpy(trunk, x, sizeof(float) * (size_t)C * s);
    /* policy head :426-507 */
    const int P = n->P, V = n->V;
    float* p = (float*)malloc(sizeof(float) * (size_t)P * s);
    float* ppool = (float*)malloc(sizeof(float) * 3 * (size_t)(P > V ? P : V));
    float* pint = (float*)malloc(sizeof(float) * 3 * (size_t)(P > V ? P : V));
    float* prob = (float*)malloc(sizeof(float) * 5 * (size_t)s);
    float pass[5], misc[15];
    conv_forward(&n->p_hd_conv, bs, x, p);
    add_spatial_biases(bs, P, p, n->p_hd_conv.b, NULL, act);
    if (n->replk) {   /* :443-471: depthwise k x k (+bias, act), then 1x1 P->P (+bias, act) */
        float* pb = (float*)malloc(sizeof(float) * (size_t)P * s);
        dwconv_forward(&n->p_dw_conv, bs, p, pb);
        add_spatial_biases(bs, P, pb, n->p_dw_conv.b, NULL, act);
        conv_forward(&n->p_pt_conv, bs, pb, p);
        add_spatial_biases(bs, P, p, n->p_pt_conv.b, NULL, act);
        free(pb);
    }
    global_pool(bs, P, p, ppool);
    fc_forward(&n->p_inter_fc, ppool, pint, act);
    add_spatial_biases(bs, P, p, pint, NULL, ACT_IDENTITY); /* :483-484 per-channel add, no activation */
    conv_forward(&n->prob_conv, bs, p, prob);
    add_spatial_biases(bs, 5, prob, n->prob_conv.b, NULL, ACT_IDENTITY);
    fc_forward(&n->pass_fc, pint, pass, ACT_IDENTITY);
    /* value head :509-555 */
    float* v = (float*)malloc(sizeof(float) * (size_t)V * s);
    float* own = (float*)malloc(sizeof(float) * (size_t)s);
    conv_forward(&n->v_hd_conv, bs, x, v);
    add_spatial_biases(bs, V, v, n->v_hd_conv.b, NULL, act);
    global_pool_value(bs, V, v, ppool);
    fc_forward(&n->v_inter_fc, ppool, pint, act);
    conv_forward(&n->v_ownership, bs, v, own);
    add_spatial_biases(bs, 1, own, n->v_ownership.b, NULL, ACT_IDENTITY);
    fc_forward(&n->v_misc, pint, misc, ACT_IDENTITY);
    /* FillOutputs :597-618 */
    memcpy(out, prob + (size_t)offset * s, sizeof(float) * (size_t)s);
    memcpy(out + s, own, sizeof(float) * (size_t)s);
    float* m = out + 2 * s;
    m[0] = pass[offset]; m[1] = misc[0]; m[2] = misc[1]; m[3] = misc[2]; m[4] = misc[3];
    m[5] = misc[8]; m[6] = misc[13]; m[7] = misc[14];
    if (all_prob) {
        memcpy(all_prob, prob, sizeof(float) * 5 * (size_t)s);
        memcpy(all_prob + 5 * s, pass, sizeof pass);
        memcpy(all_prob + 5 * s + 5, misc, sizeof misc);
    }
    free(x); free(t); free(u); free(p); free(ppool); free(pint); free(prob); free(v); free(own);
    return 0;
}

int oracle_forward(const oracle_net* n, const float* planes, int bs, int offset, float* out) {
    return oracle_forward_trace(n, planes, bs, offset, out, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------ */
/* Index work (bit-exact)                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* BatchForwardPipe::SendQueryAndWait input re-layout, batch_forward_pipe.cc:15-33:
 * n x n planes packed contiguously -> N x N canvas, top-left, zero elsewhere. */
void oracle_canvas_place(const float* planes, int channels, int n, int N, float* canvas) {
    for (int c = 0; c < channels; ++c) {
        int off_r = c * N * N, off_p = c * n * n;
        for (int idx = 0; idx < N * N; ++idx) {
            const int x = idx % N, y = idx / N;
            if (x < n && y < n) canvas[off_r++] = planes[off_p++];
            else canvas[off_r++] = 0.f;
        }
    }
}

/* Output re-layout, batch_forward_pipe.cc:48-67: N x N canvas -> n x n packed. */
void oracle_canvas_crop(const float* canvas, int n, int N, float* packed) {
    int off_r = 0, off_p = 0;
    for (int idx = 0; idx < N * N; ++idx) {
        const int x = idx % N, y = idx / N;
        if (x < n && y < n) packed[off_r++] = canvas[off_p++];
        else off_p++;
    }
}

/* Symmetry::GetSymmetry index table, game/symmetry.cc:97-123 (as summarised in SURVEY.md §8c):
 * swap x/y if symm&4, then x <- N-1-x if symm&2, then y <- N-1-y if symm&1; index = y*N + x. */
void oracle_symmetry_table(int N, int symm, int* table) {
    for (int y = 0; y < N; ++y) {
        for (int x = 0; x < N; ++x) {
            int sx = x, sy = y;
            if (symm & 4) { int tmp = sx; sx = sy; sy = tmp; }
            if (symm & 2) sx = N - 1 - sx;
            if (symm & 1) sy = N - 1 - sy;
            table[y * N + x] = sy * N + sx;
        }
    }
}
