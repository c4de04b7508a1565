// tools/de_probe.cu — probe of Blackwell's hardware decompression engine through the CUDA 12.8+ driver API
// (cuMemBatchDecompressAsync): is DEFLATE offered on this device, what is the largest single operation, and what
// throughput do batches of raw-deflate blocks of FASTQ-like text reach.  Developer tool, not part of the product:
//   nvcc -O2 -o gpurun_out/de_probe tools/de_probe.cu -lcuda -lz && gpurun_out/de_probe
#include <cuda.h>
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x) do { CUresult r__ = (x); if (r__ != CUDA_SUCCESS) { const char* s__ = nullptr; cuGetErrorString(r__, &s__); \
    printf("FAIL %s -> %d (%s)\n", #x, (int)r__, s__ ? s__ : "?"); return 1; } } while (0)

static std::vector<unsigned char> raw_deflate(const unsigned char* p, size_t n, int level) {
    z_stream z; memset(&z, 0, sizeof z);
    deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);       // raw deflate, no zlib/gzip wrapper
    std::vector<unsigned char> out(deflateBound(&z, n));
    z.next_in = const_cast<unsigned char*>(p); z.avail_in = (uInt)n; z.next_out = out.data(); z.avail_out = (uInt)out.size();
    deflate(&z, Z_FINISH);
    out.resize(z.total_out);
    deflateEnd(&z);
    return out;
}

int main(int argc, char**) {
    CK(cuInit(0));
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    CUcontext cx; CK(cuDevicePrimaryCtxRetain(&cx, dev)); CK(cuCtxSetCurrent(cx));
    int mask = 0, maxlen = 0;
    CK(cuDeviceGetAttribute(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, dev));
    CK(cuDeviceGetAttribute(&maxlen, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, dev));
    printf("decompress algorithm mask = %d (deflate %d snappy %d lz4 %d), maximum length = %d bytes\n", mask, mask & 1, (mask >> 1) & 1, (mask >> 2) & 1, maxlen);
    if (!(mask & 1)) { printf("DEFLATE not offered by this device/driver\n"); return 0; }
    // FASTQ-like text: 4-line records, 150 bp reads with random bases, constant quality
    std::string txt;
    srand(7);
    const size_t target = 256u << 20;
    while (txt.size() < target) {
        char hdr[64]; snprintf(hdr, sizeof hdr, "@s0_%zu/1\n", txt.size() / 320);
        txt += hdr;
        for (int i = 0; i < 150; i++) txt += "ACGT"[rand() & 3];
        txt += "\n+\n";
        txt.append(150, 'I');
        txt += "\n";
    }
    // alignment: BGZF payloads start 18 bytes into a member and members follow each other without padding; text
    // destinations follow each other at arbitrary offsets
    for (int mis : {0, 1, 2, 7}) {
        const size_t block = 65280, nb = 512;
        std::vector<std::vector<unsigned char>> comp(nb);
        size_t ctot = 0;
        for (size_t b = 0; b < nb; b++) { comp[b] = raw_deflate((const unsigned char*)txt.data() + b * block, block, 1); ctot += comp[b].size() + 26; }
        CUdeviceptr dsrc, ddst, dact;
        CK(cuMemAlloc(&dsrc, ctot + 64)); CK(cuMemAlloc(&ddst, nb * block + 64)); CK(cuMemAlloc(&dact, nb * 4));
        std::vector<unsigned char> flat(ctot + 64);
        std::vector<CUmemDecompressParams> ps(nb);
        size_t off = mis;
        for (size_t b = 0; b < nb; b++) {
            memcpy(flat.data() + off, comp[b].data(), comp[b].size());
            memset(&ps[b], 0, sizeof ps[b]);
            ps[b].srcNumBytes = comp[b].size(); ps[b].dstNumBytes = block; ps[b].dstActBytes = (cuuint32_t*)(dact + b * 4);
            ps[b].src = (const void*)(dsrc + off); ps[b].dst = (void*)(ddst + mis + b * block); ps[b].algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
            off += comp[b].size() + 26 - (b % 3);          // arbitrary, unaligned strides
        }
        CK(cuMemcpyHtoD(dsrc, flat.data(), ctot + 32));
        CUstream st; CK(cuStreamCreate(&st, CU_STREAM_NON_BLOCKING));
        size_t erridx = 0;
        CUresult r = cuMemBatchDecompressAsync(ps.data(), nb, 0, &erridx, st);
        CUresult r2 = r == CUDA_SUCCESS ? cuStreamSynchronize(st) : r;
        std::vector<unsigned char> back(nb * block);
        bool ok = false;
        if (r2 == CUDA_SUCCESS) { CK(cuMemcpyDtoH(back.data(), ddst + mis, nb * block)); ok = memcmp(back.data(), txt.data(), nb * block) == 0; }
        printf("misaligned by %d (src and dst): submit %d, sync %d, output %s\n", mis, (int)r, (int)r2, ok ? "identical" : "DIFFERS");
        cuMemFree(dsrc); cuMemFree(ddst); cuMemFree(dact);
    }
    if (argc > 1) return 0;                      // any argument: alignment test only
    for (size_t block : {size_t(64) << 10, size_t(1) << 20, size_t(4) << 20}) {
        if ((size_t)maxlen && block > (size_t)maxlen) { printf("block %zu > maximum length, skipped\n", block); continue; }
        const size_t nb = txt.size() / block;
        std::vector<std::vector<unsigned char>> comp(nb);
        size_t ctot = 0;
        auto t0 = std::chrono::steady_clock::now();
        for (size_t b = 0; b < nb; b++) { comp[b] = raw_deflate((const unsigned char*)txt.data() + b * block, block, 6); ctot += (comp[b].size() + 15) & ~size_t(15); }
        const double tc = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        CUdeviceptr dsrc, ddst, dact;
        CK(cuMemAlloc(&dsrc, ctot + 64)); CK(cuMemAlloc(&ddst, nb * block)); CK(cuMemAlloc(&dact, nb * 4));
        std::vector<unsigned char> flat(ctot + 64);
        std::vector<CUmemDecompressParams> ps(nb);
        size_t off = 0;
        for (size_t b = 0; b < nb; b++) {
            memcpy(flat.data() + off, comp[b].data(), comp[b].size());
            memset(&ps[b], 0, sizeof ps[b]);
            ps[b].srcNumBytes = comp[b].size(); ps[b].dstNumBytes = block; ps[b].dstActBytes = (cuuint32_t*)(dact + b * 4);
            ps[b].src = (const void*)(dsrc + off); ps[b].dst = (void*)(ddst + b * block); ps[b].algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
            off += (comp[b].size() + 15) & ~size_t(15);
        }
        CK(cuMemcpyHtoD(dsrc, flat.data(), ctot));
        CUstream st; CK(cuStreamCreate(&st, CU_STREAM_NON_BLOCKING));
        CUevent e0, e1; CK(cuEventCreate(&e0, 0)); CK(cuEventCreate(&e1, 0));
        size_t erridx = 0;
        CUresult r = cuMemBatchDecompressAsync(ps.data(), nb, 0, &erridx, st);          // warm-up
        if (r != CUDA_SUCCESS) { const char* s = nullptr; cuGetErrorString(r, &s); printf("block %zu: cuMemBatchDecompressAsync -> %d (%s), index %zu\n", block, (int)r, s ? s : "?", erridx); continue; }
        CK(cuStreamSynchronize(st));
        float best = 1e30f;
        for (int it = 0; it < 5; it++) {
            CK(cuEventRecord(e0, st));
            CK(cuMemBatchDecompressAsync(ps.data(), nb, 0, &erridx, st));
            CK(cuEventRecord(e1, st));
            CK(cuStreamSynchronize(st));
            float ms; CK(cuEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        std::vector<unsigned char> back(nb * block);
        CK(cuMemcpyDtoH(back.data(), ddst, nb * block));
        std::vector<unsigned> act(nb);
        CK(cuMemcpyDtoH(act.data(), dact, nb * 4));
        bool ok = memcmp(back.data(), txt.data(), nb * block) == 0;
        for (size_t b = 0; b < nb; b++) ok = ok && act[b] == block;
        printf("block %8zu B x %zu: ratio %.2f, host deflate %.2f s, engine %.3f ms = %.1f GB/s out (%.1f GB/s in), output %s\n", block, nb,
               (double)(nb * block) / ctot, tc, best, nb * block / best / 1e6, ctot / best / 1e6, ok ? "identical" : "DIFFERS");
        cuMemFree(dsrc); cuMemFree(ddst); cuMemFree(dact);
    }
    return 0;
}
