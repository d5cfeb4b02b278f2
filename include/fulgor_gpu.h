/*
 * fulgor_gpu.h -- C ABI of libfulgor_gpu.so: the B200 (sm_100a) implementation of Fulgor's
 * pseudoalignment hot path, as a drop-in for the three `index<ColorSets>` member functions the
 * reference's `pseudoalign` tool calls (reference include/index.hpp:39-46):
 *
 *     fetch_color_set_ids            src/ps_full_intersection.cpp:335-374   (called src/ps_utils.cpp:278)
 *     pseudoalign_full_intersection  src/ps_full_intersection.cpp:377-400   (called tools/pseudoalign.cpp:29)
 *     pseudoalign_threshold_union    src/ps_threshold_union.cpp:321-402     (called tools/pseudoalign.cpp:32)
 *
 * The reference processes one read per call; a GPU needs batches, so the boundary sits one level
 * up, at the body of `pseudoalign_worker` (tools/pseudoalign.cpp:22-52): a chunk of reads in,
 * per-read color lists out, in CSR form.
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary; no torch / C++ types in signatures.
 *   - return 0 on success, a negative errno-style code otherwise (FULGOR_GPU_E*); the message is
 *     available from fulgor_gpu_last_error() on the calling thread.
 *   - the caller owns every buffer. Host buffers may be pageable; pinned memory
 *     (fulgor_gpu_host_alloc) is recommended and is what the end-to-end numbers are measured with.
 *   - results of read i are values[off[i] .. off[i+1]), ascending: exactly the vector the reference
 *     hands to formatter.format(query_id, colors) (src/ps_utils.cpp:55,127,168).
 *   - when `cap` is too small the call returns FULGOR_GPU_E2BIG and off[n_reads] holds the required
 *     capacity (offsets are complete, values are not), so the caller can retry.
 *   - one in-flight call per handle; a handle is bound to one CUDA device.
 *   - there is NO CPU fallback: every entry point that computes fails with FULGOR_GPU_ENODEV when
 *     no CUDA device is usable.
 */
#ifndef FULGOR_GPU_H
#define FULGOR_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FULGOR_GPU_EIO (-5)      /* CUDA failure, unreadable file */
#define FULGOR_GPU_E2BIG (-7)    /* output capacity too small; off[n] = required */
#define FULGOR_GPU_ENOMEM (-12)
#define FULGOR_GPU_ENODEV (-19)  /* no usable CUDA device */
#define FULGOR_GPU_EINVAL (-22)

#define FULGOR_GPU_FULL_INTERSECTION 0 /* reference: pseudoalignment_algorithm::FULL_INTERSECTION */
#define FULGOR_GPU_THRESHOLD_UNION 1   /* reference: pseudoalignment_algorithm::THRESHOLD_UNION */

typedef struct fulgor_gpu_index fulgor_gpu_index; /* opaque: device image + per-call scratch */

/* what index<ColorSets>::{k, num_colors, num_unitigs, num_color_sets} and sshash::dictionary::{m, size}
   report (include/index.hpp:56-66) */
typedef struct fulgor_gpu_info {
    uint32_t k, m;
    uint64_t num_kmers, num_unitigs, num_color_sets;
    uint32_t num_colors;
    uint32_t type;        /* 0 = hybrid (.fur), 1 = meta (.mfur), 2 = differential (.dfur), 3 = meta-differential (.mdfur) */
    uint64_t image_bytes; /* size of the flattened device image */
    int32_t device;       /* CUDA device ordinal the handle is bound to; -1 for a host-only image */
    uint32_t pad;
} fulgor_gpu_info;

/* ---- index image: host side, no GPU needed -------------------------------------------------- */

/* Replaces essentials::load(index, path) (reference tools/pseudoalign.cpp:340): parses a
   reference-built .fur / .mfur / .dfur / .mdfur (type by file suffix, tools/util.cpp:5-19) and flattens it into one
   position-independent byte image. *image is malloc'ed by the library; release it with
   fulgor_gpu_image_free. Load errors mirror the reference's std::runtime_error texts. */
int fulgor_gpu_image_build(const char* index_path, uint8_t** image, uint64_t* image_bytes);
void fulgor_gpu_image_free(uint8_t* image);
/* header fields of an image (host pointer) */
int fulgor_gpu_image_info(const uint8_t* image, uint64_t image_bytes, fulgor_gpu_info* out);

/* ---- handles ----------------------------------------------------------------------------- */

/* build the image from the file and upload it to CUDA device `device` */
int fulgor_gpu_index_open(const char* index_path, int device, fulgor_gpu_index** out);
/* upload a host image (e.g. one received from another process) to `device` */
int fulgor_gpu_index_open_image(const uint8_t* image, uint64_t image_bytes, int device, fulgor_gpu_index** out);
/* adopt an image that is ALREADY in device memory of `device` (e.g. the receive buffer of an NCCL
   broadcast: multi-GPU replication needs one broadcast of the bytes and no other collective). The
   memory stays owned by the caller and must outlive the handle. */
int fulgor_gpu_index_adopt_device_image(const void* device_image, uint64_t image_bytes, int device, fulgor_gpu_index** out);
void fulgor_gpu_index_close(fulgor_gpu_index*);
int fulgor_gpu_index_info(const fulgor_gpu_index*, fulgor_gpu_info* out);

/* ---- the hot path, host buffers in / host buffers out (H2D and D2H inside the call) -------- */

/* Stage 1. Replaces index::fetch_color_set_ids for a batch: per read, the ascending distinct
   color-set ids of its positive k-mers; num_positive[i] (nullable) = number of positive k-mers
   (what pseudoalign_threshold_union counts, src/ps_threshold_union.cpp:327-347).
   bases: concatenated read characters; read i = bases[read_off[i] .. read_off[i+1]) -- read_off[0] need not be 0, so a
   slice of a larger batch is passed as (bases, read_off + first, n). The same holds for every entry point below. */
int fulgor_gpu_fetch_color_set_ids(fulgor_gpu_index*, const char* bases, const uint64_t* read_off, uint32_t n_reads,
                                   uint64_t* cid_off /* n_reads+1 */, uint32_t* cids, uint64_t cids_cap,
                                   uint32_t* num_positive /* n_reads, nullable */);

/* Whole path. algo = FULGOR_GPU_FULL_INTERSECTION: fetch_color_set_ids + pseudoalign_full_intersection;
   algo = FULGOR_GPU_THRESHOLD_UNION: pseudoalign_threshold_union(sequence, colors, threshold), threshold in (0,1]. */
int fulgor_gpu_pseudoalign(fulgor_gpu_index*, int algo, double threshold,
                           const char* bases, const uint64_t* read_off, uint32_t n_reads,
                           uint64_t* color_off /* n_reads+1 */, uint32_t* colors, uint64_t colors_cap);

/* Full intersection with cross-read deduplication of the color-set-id lists. Replaces the reference's `--deduplicate` mode:
   fetch_and_deduplicate_sets (tools/pseudoalign.cpp:92-226: fetch every read's list, sort the lists, keep one copy of each)
   followed by pseudoalign_worker over preprocessed_query_reader (tools/pseudoalign.cpp:39-44, src/ps_utils.cpp:307-415: one
   intersection per distinct list, written once per read id that has it). Here reads with the same list form a group over the
   WHOLE call (the per-read lists of all n_reads stay in device memory, 256 bytes per read, until the groups are known); the
   first one found is the group's representative and the only one whose intersection is computed, emitted and copied back:
     rep_of_read[i] = index of the read that represents read i (== i for a representative, and for every read without a
                      positive k-mer);
     colors of read i = colors[color_off[rep_of_read[i]] .. color_off[rep_of_read[i] + 1])  (reads that are not
                      representatives own an empty range).
   The per-read results are exactly those of fulgor_gpu_pseudoalign(FULL_INTERSECTION); like in the reference, which read of a
   group does the work is not deterministic, and threshold-union cannot be deduplicated (tools/pseudoalign.cpp:283-289). */
int fulgor_gpu_pseudoalign_dedup(fulgor_gpu_index*, const char* bases, const uint64_t* read_off, uint32_t n_reads,
                                 uint32_t* rep_of_read /* n_reads */,
                                 uint64_t* color_off /* n_reads+1 */, uint32_t* colors, uint64_t colors_cap);

/* ---- compact forms of the same path: fewer bytes over PCIe per read --------------------------------------------------
   The reference's worker owns std::string reads and std::vector<uint32_t> results (tools/pseudoalign.cpp:22-52); a GPU worker is
   bound by the copies of exactly those, so both have a compact form here. Results are identical to fulgor_gpu_pseudoalign's.

   PACKED READS: what the reference's own k-mer type holds (external/sshash/include/kmer.hpp:199: 2 bits per base,
   code = (c >> 1) & 3, i.e. A 0, C 1, T 2, G 3, either case), 16 bases per 32-bit word, base p of a read in bits [2 (p % 16), +2) of
   its word p / 16; EVERY READ STARTS ON A WORD: read i occupies words [W_i, W_i + ceil(len_i / 16)), W_i = sum of the earlier
   reads' word counts; unused bits of a read's last word are ignored.
     read_len[i]    = len_i (< 2^31), with FULGOR_GPU_READ_HAS_INVALID set when the read holds a character other than ACGTacgt
                      (such characters invalidate every k-mer over them, kmer.hpp:214-224,258-260; their codes are ignored);
     invalid_pos[]  = ascending positions 16 * W_i + p of those characters, n_invalid of them (NULL / 0 when there are none).
   fulgor_gpu_pack_reads produces this form from ASCII (host threads; what a FASTQ tokeniser would emit directly).

   BITMAP RESULTS: instead of CSR lists, one row of ceil(num_colors / 32) 32-bit words per read, bit c % 32 of word c / 32 set
   iff color c is in the read's result -- the set the reference writes as a list (src/ps_utils.cpp:127-135). 4 bytes per read
   for up to 32 colors, 572 bytes for 4,546 colors (a full-intersection list there averages kilobytes). */
#define FULGOR_GPU_READ_HAS_INVALID 0x80000000u

/* ASCII reads -> packed reads with up to `threads` host threads (<= 0: all). Returns FULGOR_GPU_E2BIG when words_cap or
   invalid_cap is too small; *n_words / *n_invalid then hold the required capacities. No GPU needed. */
int fulgor_gpu_pack_reads(const char* bases, const uint64_t* read_off, uint32_t n_reads,
                          uint32_t* words, uint64_t words_cap, uint32_t* read_len /* n_reads */,
                          uint64_t* invalid_pos, uint64_t invalid_cap, uint64_t* n_words, uint64_t* n_invalid, int threads);

/* fulgor_gpu_pseudoalign on packed reads, CSR lists out */
int fulgor_gpu_pseudoalign_packed(fulgor_gpu_index*, int algo, double threshold,
                                  const uint32_t* words, const uint32_t* read_len, uint32_t n_reads,
                                  const uint64_t* invalid_pos, uint64_t n_invalid,
                                  uint64_t* color_off /* n_reads+1 */, uint32_t* colors, uint64_t colors_cap);
/* fulgor_gpu_pseudoalign with bitmap rows out: bitmaps = n_reads * ceil(num_colors / 32) words */
int fulgor_gpu_pseudoalign_bitmaps(fulgor_gpu_index*, int algo, double threshold,
                                   const char* bases, const uint64_t* read_off, uint32_t n_reads, uint32_t* bitmaps);
/* packed reads in, bitmap rows out */
int fulgor_gpu_pseudoalign_packed_bitmaps(fulgor_gpu_index*, int algo, double threshold,
                                          const uint32_t* words, const uint32_t* read_len, uint32_t n_reads,
                                          const uint64_t* invalid_pos, uint64_t n_invalid, uint32_t* bitmaps);

/* ---- the per-k-mer tools that share the lookup kernel (reference tools/kmer_conservation.cpp, tools/kmer_matches.cpp) ---- */

/* Replaces index::kmer_conservation (src/kmer_conservation.cpp:7-54, called tools/kmer_conservation.cpp:26) for a batch: per
   read the maximal runs of consecutive positive k-mers with one color-set id, as kmer_conservation_triple
   {start_pos_in_query, num_kmers, color_set_id} (include/util.hpp:74-78), in query order. Read i owns triples
   [triple_off[i], triple_off[i+1]); triple t is triples[3t .. 3t+2]. triples_cap counts triples. */
int fulgor_gpu_kmer_conservation(fulgor_gpu_index*, const char* bases, const uint64_t* read_off, uint32_t n_reads,
                                 uint64_t* triple_off /* n_reads+1 */, uint32_t* triples, uint64_t triples_cap);

/* Replaces index::kmer_matches (src/kmer_matches.cpp:7-30, called tools/kmer_matches.cpp:25) for a batch:
     positive k-mers: read i owns the 32-bit words [word_off[i], word_off[i+1]) of positive_words, ceil(num_kmers / 32) of
                      them; bit j (LSB first) = k-mer j of the read is in the index (positive_kmers_in_sequence);
     counts:          n_reads x num_colors, counts[i * num_colors + c] = positive k-mers of read i whose color set contains c.
   Reads shorter than k have no k-mers: no words and zero counts (the reference returns early there and its caller prints
   whatever the previous read left in the buffers, src/kmer_matches.cpp:11). Needs the decoded color-set table. */
int fulgor_gpu_kmer_matches(fulgor_gpu_index*, const char* bases, const uint64_t* read_off, uint32_t n_reads,
                            uint64_t* word_off /* n_reads+1 */, uint32_t* positive_words, uint64_t words_cap,
                            uint32_t* counts /* n_reads * num_colors */);

/* ---- the same path on DEVICE-resident inputs (kernel-only timing, pipelines that keep reads on
        the GPU). All pointers are device pointers on the handle's device; offsets are CSR like above.
        *total_out (host) receives off[n_reads]. Runs on the handle's stream and synchronises it. */
int fulgor_gpu_pseudoalign_device(fulgor_gpu_index*, int algo, double threshold,
                                  const char* d_bases, const uint64_t* d_read_off, uint32_t n_reads, uint64_t read_off_base,
                                  uint64_t* d_color_off, uint32_t* d_colors, uint64_t colors_cap, uint64_t* total_out);

/* the same on packed reads resident on the device (d_words, d_read_len, d_invalid_pos: device pointers; positions relative
   to d_words[0]). result_bitmaps != 0: d_colors receives n_reads bitmap rows (colors_cap counts 32-bit words, d_color_off is
   not written, *total_out = words written); else CSR lists like fulgor_gpu_pseudoalign_device. */
int fulgor_gpu_pseudoalign_packed_device(fulgor_gpu_index*, int algo, double threshold,
                                         const uint32_t* d_words, const uint32_t* d_read_len, uint32_t n_reads,
                                         const uint64_t* d_invalid_pos, uint32_t n_invalid, int result_bitmaps,
                                         uint64_t* d_color_off, uint32_t* d_colors, uint64_t colors_cap, uint64_t* total_out);

/* device time (ms, CUDA events on the handle's stream) of the kernels of the last *_device call:
   [0] k-mer lookup (+ fused small-color-count intersect/union), [1] color-set kernel (0 when fused),
   [2] offsets scan + emit. Returns the number of kernel launches of that call. */
int fulgor_gpu_last_kernel_times(const fulgor_gpu_index*, float ms[3]);

/* ---- utilities ----------------------------------------------------------------------------- */
int fulgor_gpu_device_count(void);
void* fulgor_gpu_host_alloc(uint64_t bytes); /* pinned host memory (cudaHostAlloc); NULL on failure */
/* Narrows the calling thread's CPU affinity to the CPUs next to `device` (sysfs local_cpulist of its PCI function), so that the
   pinned buffers the thread allocates afterwards and the copies it issues stay on the device's NUMA node: with one process or
   thread per GPU on a multi-socket host this keeps N host->device streams from crossing the socket interconnect. Returns the
   number of CPUs bound, 0 when there is nothing to narrow or the topology is unknown (never an error). Threads created
   afterwards inherit the mask. */
int fulgor_gpu_bind_host_thread(int device);
void fulgor_gpu_host_free(void*);
const char* fulgor_gpu_last_error(void);     /* thread-local */
const char* fulgor_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
