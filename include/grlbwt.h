/*
 * grlbwt.h -- C entry points of the host side (libgrlbwt.so): the whole BCR BWT construction
 * (device parse phase through include/grlgpu.h + host induction phase + .rl_bwt writer).
 * Mirrors grl_bwt_algo<sym_type,false>() of the reference (include/grl_bwt.hpp:23-79) for
 * callers that cannot include C++ templates (tests, Python via ctypes).
 */
#ifndef GRLBWT_H
#define GRLBWT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t n_runs;
    uint64_t sb, fb;          /* header of the .rl_bwt (bwt_io.h:377-382) */
    uint64_t* syms;           /* n_runs, malloc'ed; release with grlbwt_free_result */
    uint64_t* lens;
    uint64_t n_rounds;
    double h2d_ms;            /* host -> device copy of the text */
    double par_phase_ms;      /* device parse rounds + fetch of the level artefacts */
    double ind_phase_ms;      /* host induction */
    double device_ms;         /* sum of CUDA-event round times */
    uint64_t algorithmic_bytes; /* sum of B_r over the rounds */
    uint64_t induced_on_device; /* 1: the induction phase ran on the GPU (ind_phase_ms is its time); 0: on the host threads */
} grlbwt_result_t;

/* BWT of a collection held in host memory; status codes of grlgpu.h (or -100 for host-side errors) */
int grlbwt_build(const void* text, uint64_t n_syms, int sym_bytes, int device, int n_threads, int verbose, grlbwt_result_t* out);
void grlbwt_free_result(grlbwt_result_t* r);
/* the same over several GPUs: one rank (host thread) per entry of `devices`, shards of whole strings, partitioned global
 * dictionary (include/grlgpu.h, multi-GPU rounds). A device may be listed more than once (ranks then share it: this is how
 * the N > 1 path runs on a 1-GPU box). comm_kind: 0 = in-process peer copies when every pair of GPUs has peer access (or ranks share a
 * GPU), else NCCL; 1 = in-process peer copies; 2 = NCCL. The bytes of the result do not depend on the number of ranks. */
int grlbwt_build_mg(const void* text, uint64_t n_syms, int sym_bytes, const int* devices, int n_ranks, int n_threads, int comm_kind, int verbose,
                    grlbwt_result_t* out);
/* the same construction with the level-0 BWT written into CALLER-OWNED arrays of cap_runs 32-bit symbols / 32-bit lengths
 * (out->syms / out->lens stay NULL). With pinned buffers and the induction on the device the runs arrive by DMA, without a
 * first touch of fresh host pages; a production caller reuses the buffers across collections. GRLGPU_ERR_LIMIT when the BWT
 * has more than cap_runs runs or does not fit 32 bits. */
int grlbwt_build_to(const void* text, uint64_t n_syms, int sym_bytes, const int* devices, int n_ranks, int n_threads, int comm_kind, uint32_t* out_syms,
                    uint32_t* out_lens, uint64_t cap_runs, grlbwt_result_t* out);
/* the same construction with the level-0 BWT delivered as the IMAGE OF THE .rl_bwt FILE the reference writes (main.cpp:146-152):
 * [sb u64][fb u64] then n_runs records of sb symbol bytes + fb length bytes. With the induction on the device the records are packed
 * there, so sb + fb bytes per run cross PCIe instead of 8; sha256(image) == sha256 of the reference's output file. *image_bytes =
 * 16 + n_runs * (sb + fb); GRLGPU_ERR_LIMIT when cap_bytes is smaller. */
int grlbwt_build_packed(const void* text, uint64_t n_syms, int sym_bytes, const int* devices, int n_ranks, int n_threads, int comm_kind, void* out_image,
                        uint64_t cap_bytes, uint64_t* image_bytes, grlbwt_result_t* out);
/* per-round digests of the calling thread's last build (9 values per round: tot_phrases, pre-BWT runs, parse length, distinct
 * phrases, dictionary symbols, and the four sums of grlgpu_level_checksum added over the ranks); returns the number of rounds */
uint64_t grlbwt_last_digests(uint64_t* out, uint64_t cap_rounds);
uint64_t grlbwt_last_exchange_bytes(void);   /* bulk bytes exchanged between the ranks during the last build */
const char* grlbwt_last_comm(void);          /* exchange backend of the last build */

/* same as the CLI: TEXT file -> .rl_bwt file (main.cpp:98-154 + grl_bwt.hpp:23-79) */
int grlbwt_build_file(const char* input_file, const char* output_file, int sym_bytes, int device, int n_threads, int verbose);

const char* grlbwt_last_error(void);

/* self test of the host induction phase alone (no device): levels given as parallel arrays (u64 symbols; the
 * multi-threaded path narrows them to 32 bits),
 * level 0 = round 1; fills out->syms / out->lens / out->n_runs */
int grlbwt_selftest_induce(int n_levels, const uint64_t* alphabet, const uint64_t* tot, const uint64_t* const* rule_l, const uint64_t* const* rule_r,
                           const uint8_t* const* has_hocc, const uint64_t* n_pre, const uint64_t* const* pre_sym, const uint64_t* const* pre_len,
                           const uint64_t* final_parse, uint64_t n_strings, int n_threads /* 0 = sequential 64-bit path */, grlbwt_result_t* out);

/* self test of the multi-GPU sharding rule alone (no device): contiguous ranges of whole strings balanced by symbol count;
 * bounds_out receives n_shards + 1 cell offsets (n_shards <= n_ranks: fewer when the collection has fewer strings) */
int grlbwt_selftest_shard_bounds(const void* text, uint64_t n_syms, int sym_bytes, int n_ranks, uint64_t* bounds_out, int* n_shards);

/* self test of the .rl_bwt writer alone (no device; format of include/bwt_io.h:377-382,448-490): runs given as u64
 * symbols / lengths; narrow != 0 routes through the 32-bit-symbol instantiation the multi-threaded host uses */
int grlbwt_selftest_write(const char* path, const uint64_t* syms, const uint64_t* lens, uint64_t n_runs, uint64_t sb, uint64_t fb, int narrow);
/* the same records packed into memory ([sb][fb] + records: the image grlbwt_build_packed returns); narrow != 0 packs from 32-bit
 * symbols and lengths */
int grlbwt_selftest_pack(void* out_image, uint64_t cap_bytes, const uint64_t* syms, const uint64_t* lens, uint64_t n_runs, uint64_t sb, uint64_t fb, int narrow);

#ifdef __cplusplus
}
#endif
#endif
