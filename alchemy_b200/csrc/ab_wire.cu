// ab_wire.cu -- a wire / on-disk format for batches of ring elements, ciphertexts and hints.
// The reference keeps keys and hints as in-memory Haskell `Dynamic`s (Crypto/Alchemy/Interpreter/KeysHints.hs:36-39) and
// has no serialisation; sharding ciphertexts over processes or boxes needs one.  Format: a 136-byte little-endian
// header (include/alchemy_b200.h `ab_wire_header`) followed by one uint32 per residue in the device layout
// [batch][npoly][nlimbs][n] -- residues are below 2^31, so the payload is half the size of the int64
// tensors it replaces on NVLink / PCIe / disk.  The header carries a position-weighted checksum of its own fields and of the payload;
// packing refuses values that are not 31-bit residues, unpacking recomputes the checksum and refuses truncated or corrupted input.
#include <algorithm>
#include <cstddef>
#include <cstring>

#include "ab_internal.h"

static const uint32_t kWireMagic = 0x57324241u;      // "AB2W"
static const uint16_t kWireVersion = 2;               // 2: position-weighted checksum that also covers the header fields

static int64_t wire_count(const ab_wire_header *h) { return h->batch * h->npoly * h->nlimbs * h->n; }
static bool wire_sane(const ab_wire_header *h) {
    return h->magic == kWireMagic && h->version == kWireVersion && h->kind >= 1 && h->kind <= 4 && h->m >= 1 && h->n >= 1 && h->batch >= 0 &&
           h->npoly >= 1 && h->nlimbs >= 1 && h->nlimbs <= AB_MAX_LIMBS && h->n < (1ll << 24) && h->batch < (1ll << 40) && h->npoly < (1 << 20);
}
// checksum = sum_j (j+1) * K * w_j  +  sum_i (i+1) * v_i   (mod 2^64): w_j the sixteen 64-bit little-endian words of the header before
// the checksum field, v_i the payload words.  Position weights catch reordered / transposed words, the header term a flipped k, l, enc,
// modulus or size.  (The test suite restates it with numpy.)
static const uint64_t kWireK = 0x9E3779B97F4A7C15ull;
static uint64_t wire_header_sum(const ab_wire_header *h) {
    static_assert(sizeof(ab_wire_header) == 136 && offsetof(ab_wire_header, checksum) == 128, "wire header layout");
    uint64_t w[16], s = 0;
    memcpy(w, h, 128);
    for (int j = 0; j < 16; ++j) s += (uint64_t)(j + 1) * kWireK * w[j];
    return s;
}

extern "C" int64_t ab_wire_bytes(const ab_wire_header *h) {
    if (!h || h->batch < 0 || h->npoly < 1 || h->nlimbs < 1 || h->n < 1) return -1;
    return (int64_t)sizeof(ab_wire_header) + 4 * wire_count(h);
}

// st[0] += sum (i+1) v_i;  st[1] |= bits of residues that do not fit 31 bits (negative or >= 2^31: not a residue of this backend)
__global__ void __launch_bounds__(256) k_wire_pack(uint32_t *__restrict__ out, const int64_t *__restrict__ in, int64_t count, unsigned long long *st) {
    unsigned long long s = 0, bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t x = in[i];
        const uint32_t v = (uint32_t)x;
        bad |= (unsigned long long)x >> 31;
        out[i] = v;
        s += (unsigned long long)(i + 1) * v;
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); bad |= __shfl_xor_sync(0xffffffffu, bad, o); }
    if ((threadIdx.x & 31) == 0) {
        if (s) atomicAdd(st, s);
        if (bad) atomicOr(st + 1, bad);
    }
}
__global__ void __launch_bounds__(256) k_wire_unpack(int64_t *__restrict__ out, const uint32_t *__restrict__ in, int64_t count, unsigned long long *st) {
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = in[i];
        out[i] = (int64_t)v;
        s += (unsigned long long)(i + 1) * v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(st, s);
}
static unsigned wire_grid(int64_t count) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((count + 255) / 256, 148 * 16)); }

// two device words (checksum accumulator, range flag), released on every exit path
struct WireStat {
    unsigned long long *d = nullptr;
    cudaStream_t st;
    explicit WireStat(cudaStream_t s) : st(s) {
        AB_CHECK_CUDA(cudaMallocAsync((void **)&d, 16, st));
        if (cudaMemsetAsync(d, 0, 16, st) != cudaSuccess) { cudaFreeAsync(d, st); d = nullptr; throw ab_error{AB_ERR_CUDA, "cudaMemsetAsync failed"}; }
    }
    ~WireStat() { if (d) cudaFreeAsync(d, st); }
    void fetch(unsigned long long out[2]) {
        AB_CHECK_CUDA(cudaMemcpyAsync(out, d, 16, cudaMemcpyDeviceToHost, st));
        AB_CHECK_CUDA(cudaStreamSynchronize(st));
    }
};

extern "C" int ab_wire_pack_dev(void *wire_dev, const ab_wire_header *h, const int64_t *data_dev, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(wire_dev && h && (data_dev || wire_count(h) == 0), AB_ERR_ARG, "null argument");
        ab_wire_header hd = *h;
        hd.magic = kWireMagic; hd.version = kWireVersion; hd.checksum = 0;
        AB_REQUIRE(wire_sane(&hd), AB_ERR_ARG, "bad wire header");
        AB_REQUIRE(((uintptr_t)wire_dev & 7) == 0, AB_ERR_ARG, "wire buffer must be 8-byte aligned");
        cudaStream_t st = (cudaStream_t)stream;
        const int64_t count = wire_count(&hd);
        unsigned long long got[2] = {0, 0};
        if (count > 0) {
            WireStat stat(st);
            k_wire_pack<<<wire_grid(count), 256, 0, st>>>(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(wire_dev) + sizeof(ab_wire_header)), data_dev, count, stat.d);
            ab_count_launch();
            ab_check_launch("wire_pack");
            stat.fetch(got);
            AB_REQUIRE(got[1] == 0, AB_ERR_ARG, "wire pack: a residue is negative or >= 2^31 (not in [0, q))");
        }
        hd.checksum = wire_header_sum(&hd) + got[0];
        AB_CHECK_CUDA(cudaMemcpyAsync(wire_dev, &hd, sizeof(hd), cudaMemcpyHostToDevice, st));
        AB_CHECK_CUDA(cudaStreamSynchronize(st));
    });
}

extern "C" int ab_wire_read_header(ab_wire_header *h_out, const void *wire_dev, int64_t wire_bytes) {
    return ab_guard([&] {
        AB_REQUIRE(h_out && wire_dev, AB_ERR_ARG, "null argument");
        AB_REQUIRE(wire_bytes >= (int64_t)sizeof(ab_wire_header), AB_ERR_ARG, "wire buffer shorter than a header");
        AB_CHECK_CUDA(cudaMemcpy(h_out, wire_dev, sizeof(ab_wire_header), cudaMemcpyDeviceToHost));
        AB_REQUIRE(wire_sane(h_out), AB_ERR_ARG, "not an alchemy_b200 wire buffer (magic / version / sizes)");
        AB_REQUIRE(ab_wire_bytes(h_out) == wire_bytes, AB_ERR_ARG, "wire buffer truncated or oversized for its header");
    });
}

extern "C" int ab_wire_unpack_dev(int64_t *data_dev, const void *wire_dev, int64_t wire_bytes, void *stream) {
    return ab_guard([&] {
        ab_wire_header h;
        int rc = ab_wire_read_header(&h, wire_dev, wire_bytes);
        if (rc != AB_OK) throw ab_error{rc, ab_last_error()};
        const int64_t count = wire_count(&h);
        AB_REQUIRE(data_dev || count == 0, AB_ERR_ARG, "null argument");
        cudaStream_t st = (cudaStream_t)stream;
        unsigned long long got[2] = {0, 0};
        if (count > 0) {
            WireStat stat(st);
            k_wire_unpack<<<wire_grid(count), 256, 0, st>>>(data_dev, reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(wire_dev) + sizeof(ab_wire_header)), count, stat.d);
            ab_count_launch();
            ab_check_launch("wire_unpack");
            stat.fetch(got);
        }
        const uint64_t want = h.checksum;
        h.checksum = 0;
        AB_REQUIRE(wire_header_sum(&h) + got[0] == want, AB_ERR_ARG, "wire checksum mismatch: header or payload corrupted");
    });
}
