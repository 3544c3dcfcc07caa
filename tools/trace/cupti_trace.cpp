// dev aid: a kernel/memcpy timeline without nsys (not installed in this image).  Built as a shared library and
// loaded by the CUDA driver through CUDA_INJECTION64_PATH; CUPTI activity records (concurrent-kernel, memcpy,
// memset) are written as CSV to $QPZK_TRACE_OUT at exit:
//     kind,name,stream,start_ns,end_ns,grid,block,regs,smem_dyn,graph_node
// tools/trace/summarize.py turns the CSV into per-kernel busy time, concurrency and gaps per stream.
//   g++ -O2 -shared -fPIC -I/usr/local/cuda/include -o tools/trace/libqpzktrace.so tools/trace/cupti_trace.cpp -L/usr/local/cuda/lib64 -lcupti
#include <cupti.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace {
struct Rec { char kind; std::string name; uint32_t stream; uint64_t start, end; uint32_t gx, gy, gz, bx, regs, smem; uint64_t node; };
std::vector<Rec> g_recs;
std::mutex g_mu;
uint64_t g_dropped = 0;

void CUPTIAPI buf_requested(uint8_t** buffer, size_t* size, size_t* maxNumRecords) {
    *size = 8u << 20;
    *buffer = (uint8_t*)aligned_alloc(8, *size);
    *maxNumRecords = 0;
}
void CUPTIAPI buf_completed(CUcontext, uint32_t, uint8_t* buffer, size_t, size_t validSize) {
    CUpti_Activity* r = nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    while (cuptiActivityGetNextRecord(buffer, validSize, &r) == CUPTI_SUCCESS) {
        if (r->kind == CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL || r->kind == CUPTI_ACTIVITY_KIND_KERNEL) {
            auto* k = (CUpti_ActivityKernel9*)r;
            g_recs.push_back({'K', k->name ? k->name : "?", k->streamId, k->start, k->end, (uint32_t)k->gridX, (uint32_t)k->gridY,
                              (uint32_t)k->gridZ, (uint32_t)k->blockX, k->registersPerThread, (uint32_t)k->dynamicSharedMemory, k->graphNodeId});
        } else if (r->kind == CUPTI_ACTIVITY_KIND_MEMCPY) {
            auto* m = (CUpti_ActivityMemcpy6*)r;
            char nm[64]; snprintf(nm, sizeof nm, "memcpy kind %u bytes %llu", (unsigned)m->copyKind, (unsigned long long)m->bytes);
            g_recs.push_back({'C', nm, m->streamId, m->start, m->end, 0, 0, 0, 0, 0, 0, m->graphNodeId});
        } else if (r->kind == CUPTI_ACTIVITY_KIND_MEMSET) {
            auto* m = (CUpti_ActivityMemset4*)r;
            char nm[64]; snprintf(nm, sizeof nm, "memset bytes %llu", (unsigned long long)m->bytes);
            g_recs.push_back({'S', nm, m->streamId, m->start, m->end, 0, 0, 0, 0, 0, 0, m->graphNodeId});
        }
    }
    size_t dropped = 0;
    cuptiActivityGetNumDroppedRecords(nullptr, 0, &dropped);
    g_dropped += dropped;
    free(buffer);
}
void at_exit() {
    cuptiActivityFlushAll(1);
    const char* out = getenv("QPZK_TRACE_OUT");
    FILE* f = fopen(out ? out : "qpzk_trace.csv", "w");
    if (!f) return;
    fprintf(f, "kind,name,stream,start_ns,end_ns,grid,block,regs,smem_dyn,graph_node\n");
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& r : g_recs) {
        std::string n = r.name;
        for (auto& ch : n) if (ch == ',') ch = ';';
        fprintf(f, "%c,%s,%u,%llu,%llu,%ux%ux%u,%u,%u,%u,%llu\n", r.kind, n.c_str(), r.stream, (unsigned long long)r.start,
                (unsigned long long)r.end, r.gx, r.gy, r.gz, r.bx, r.regs, r.smem, (unsigned long long)r.node);
    }
    fprintf(f, "#dropped,%llu\n", (unsigned long long)g_dropped);
    fclose(f);
}
}  // namespace

extern "C" int InitializeInjection(void) {
    cuptiActivityRegisterCallbacks(buf_requested, buf_completed);
    cuptiActivityEnable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL);
    cuptiActivityEnable(CUPTI_ACTIVITY_KIND_MEMCPY);
    cuptiActivityEnable(CUPTI_ACTIVITY_KIND_MEMSET);
    atexit(at_exit);
    return 1;
}
