/* simt_emu.cpp — TEST INFRASTRUCTURE: scheduler of the lock-step SIMT emulator (see simt_emu.h). */
#include "simt_emu.h"

namespace simt {

Fiber* cur = nullptr;
Dim3 g_blockIdx, g_blockDim, g_gridDim;
void* g_dynsmem = nullptr;

static ucontext_t sched_ctx;
static const std::function<void()>* g_body = nullptr;
static const size_t STACK = 256 * 1024;

static void trampoline() {
    (*g_body)();
    cur->state = ST_DONE;
    swapcontext(&cur->ctx, &sched_ctx);
}

uint64_t collective(int op, uint64_t payload, int arg) {
    Fiber* f = cur;
    f->op = op;
    f->payload = payload;
    f->arg = arg;
    f->state = ST_WAIT_WARP;
    swapcontext(&f->ctx, &sched_ctx);
    return f->result;
}

void block_barrier() {
    Fiber* f = cur;
    f->state = ST_WAIT_BLOCK;
    swapcontext(&f->ctx, &sched_ctx);
}

static void resolve_warp(Fiber** lanes, int n) {
    /* lanes[i] == nullptr or DONE lanes do not take part; reading from them returns the caller's own value */
    int op = -1;
    for (int i = 0; i < n; i++)
        if (lanes[i] && lanes[i]->state == ST_WAIT_WARP) {
            if (op < 0) op = lanes[i]->op;
            else if (op != lanes[i]->op) {
                fprintf(stderr, "simt-emu: divergent collectives in one warp (%d vs %d)\n", op, lanes[i]->op);
                abort();
            }
        }
    uint64_t ballot = 0;
    if (op == OP_BALLOT)
        for (int i = 0; i < n; i++)
            if (lanes[i] && lanes[i]->state == ST_WAIT_WARP && lanes[i]->payload) ballot |= (1ull << i);
    for (int i = 0; i < n; i++) {
        Fiber* f = lanes[i];
        if (!f || f->state != ST_WAIT_WARP) continue;
        int src = i;
        switch (op) {
            case OP_SHFL_IDX: src = f->arg & 31; break;
            case OP_SHFL_UP: src = i - f->arg; if (src < 0) src = i; break;
            case OP_SHFL_DOWN: src = i + f->arg; if (src > 31) src = i; break;
            case OP_SHFL_XOR: src = i ^ f->arg; break;
            default: break;
        }
        if (op == OP_BALLOT) f->result = ballot;
        else if (op == OP_SYNCWARP) f->result = 0;
        else {
            Fiber* s = (src >= 0 && src < n) ? lanes[src] : nullptr;
            f->result = (s && s->state == ST_WAIT_WARP) ? s->payload : f->payload;
        }
    }
    for (int i = 0; i < n; i++)
        if (lanes[i] && lanes[i]->state == ST_WAIT_WARP) lanes[i]->state = ST_READY;
}

void launch(Dim3 grid, Dim3 block, const std::function<void()>& body, size_t dyn_smem) {
    g_body = &body;
    void* dyn = dyn_smem ? calloc(1, dyn_smem) : nullptr;
    g_dynsmem = dyn;
    g_gridDim = grid;
    g_blockDim = block;
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    std::vector<Fiber> fibers(nthreads);
    for (int t = 0; t < nthreads; t++) fibers[t].stack = (char*)malloc(STACK);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx = Dim3(bx, by, bz);
                for (int t = 0; t < nthreads; t++) {
                    Fiber& f = fibers[t];
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = STACK;
                    f.ctx.uc_link = &sched_ctx;
                    makecontext(&f.ctx, trampoline, 0);
                    f.state = ST_READY;
                    f.tid = Dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                    f.warp = t / 32;
                    f.lane = t % 32;
                }
                for (;;) {
                    bool ran = false;
                    for (int t = 0; t < nthreads; t++) {
                        if (fibers[t].state != ST_READY) continue;
                        cur = &fibers[t];
                        swapcontext(&sched_ctx, &fibers[t].ctx);
                        ran = true;
                    }
                    /* resolve warp collectives whose live lanes have all arrived */
                    bool released = false;
                    for (int w = 0; w < nwarps; w++) {
                        Fiber* lanes[32];
                        int live = 0, waiting = 0;
                        for (int l = 0; l < 32; l++) {
                            int t = w * 32 + l;
                            lanes[l] = t < nthreads ? &fibers[t] : nullptr;
                            if (!lanes[l] || lanes[l]->state == ST_DONE) continue;
                            live++;
                            if (lanes[l]->state == ST_WAIT_WARP) waiting++;
                        }
                        if (live > 0 && waiting == live) {
                            resolve_warp(lanes, 32);
                            released = true;
                        }
                    }
                    int live = 0, at_bar = 0;
                    for (int t = 0; t < nthreads; t++) {
                        if (fibers[t].state == ST_DONE) continue;
                        live++;
                        if (fibers[t].state == ST_WAIT_BLOCK) at_bar++;
                    }
                    if (live == 0) break;
                    if (at_bar == live) {
                        for (int t = 0; t < nthreads; t++)
                            if (fibers[t].state == ST_WAIT_BLOCK) fibers[t].state = ST_READY;
                        released = true;
                    }
                    if (!ran && !released) {
                        fprintf(stderr, "simt-emu: deadlock (live %d, at barrier %d)\n", live, at_bar);
                        abort();
                    }
                }
            }
    for (int t = 0; t < nthreads; t++) free(fibers[t].stack);
    free(dyn);
    g_dynsmem = nullptr;
    cur = nullptr;
}

} // namespace simt
