// enc_img.cuh -- per-image view of the encoder workspace, shared by the CUDA kernels and by
// the host-compiled test harness (tests/hostemu).  Every array of an image sits in its own
// slot with a zero guard band on both sides: the reference reads a little outside several of
// its heap blocks and reads malloc'd memory before writing it, and the canonical oracle
// defines those reads as 0.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NHW_HD __host__ __device__ __forceinline__
#define NHW_HDN __host__ __device__ inline
#else
#define NHW_HD inline
#define NHW_HDN inline
#endif

#define NHW_GUARD_S 4096                       // guard, in int16 elements (8 KiB) on each side
#define NHW_GUARD_B 4096                       // guard, in bytes, for byte arrays
#define NHW_Y_SLOT (512 * 512 + 2 * NHW_GUARD_S)
#define NHW_C_SLOT (256 * 256 + 2 * NHW_GUARD_S)
#define NHW_Q_SLOT (128 * 128 + 2 * NHW_GUARD_S)

// capacities of the side-channel lists (bytes).  The reference's own buffers are smaller in
// places (e.g. `highres`, 24577 bytes, encoder/nhw_encoder.c:1424) and it would write out of
// bounds beyond them; we flag NHW_ERR_OVERFLOW instead.
#define NHW_CAP_TREE1 (24577 + 15)
#define NHW_CAP_LIST 65536 + 64
#define NHW_CAP_EXW 16384
#define NHW_SCAN_BYTES 393216                  // 6*IM_SIZE, encoder/nhw_encoder.c:2108

// Scalar results that end up in the .nhw header (encoder/codec.h:125-181).
struct EncHdr {
	int32_t status;
	int32_t quality;
	int32_t res_high;         // setup->RES_HIGH after Y_highres_compression
	int32_t wavelet_type;
	int32_t size_tree1, size_tree2, size_data1, size_data2, tree_end, exw_Y_end;
	int32_t res1_len, res1_bit_len, res1_word_len;
	int32_t res3_len, res3_bit_len, res3_word_len;
	int32_t res4_len;
	int32_t res5_len, res5_bit_len, res5_word_len;
	int32_t res6_len, res6_bit_len, res6_word_len, char_res1_len, qsetting3_len;
	int32_t select1, select2, highres_comp_len, end_ch_res, highres_mem_len;
	int32_t stream_len;
	int32_t res_low;          // setup->RES_LOW chosen by the luma LL coder (0/1/2)
	int32_t y_res_comp;       // length of the luma part of the LL code
	int32_t exw_y_len, exw_u_len, exw_v_len;   // exw_Y pieces (spliced with 0,0 separators)
	int32_t pad[3];
};

struct EncImg {
	// luma working planes (512x512, stride 512) and 256x256 side planes
	int16_t *proc;       // im_process
	int16_t *jpeg;       // im_jpeg
	int16_t *aux;        // scratch plane
	int16_t *ll1;        // res256
	int16_t *ll2s;       // resIII
	// chroma working planes (256x256, stride 256) of the component being coded
	int16_t *cproc;
	int16_t *cjpeg;
	int16_t *caux;
	int16_t *cll1;       // chroma res256 (128x128)
	int16_t *cll2s;      // chroma resIII (128x128)
	// byte streams / lists
	uint8_t *scan;       // im_nhw, 393216 bytes
	uint8_t *tree1;      // LL bytes: Y 16384 + U 4096 + V 4096 (+1)
	uint8_t *ch_res;     // un-truncated LL2 bytes (E11), 16384
	uint8_t *llcode;     // LL code: luma part then chroma part (the file's ch_res section)
	uint8_t *exw;        // exw_Y (luma part)
	uint8_t *exw_uv;     // exw_Y entries of U (first 16384 bytes) and V (next 16384), spliced in by the writer
	uint8_t *res1, *res1_bit, *res1_word;
	uint8_t *res3, *res3_bit, *res3_word;
	uint8_t *res4;
	uint8_t *res5, *res5_bit, *res5_word;
	uint8_t *res6, *res6_bit, *res6_word;   // q22/q23 side channel (enc_hq.cuh)
	uint16_t *char_res1;
	uint32_t *qsetting3;
	int16_t *hq_qs, *hq_fo, *hq_band;       // kept first pass (256x512), LL1 copy (256x256), rebuilt LH1 (256x256)
	uint8_t *hq_tag;                        // 256x512 comparison tags
	uint8_t *tmp1, *tmp2, *tmp3;   // list scratch (highres / ch_comp / scan_run)
	uint16_t *highres_mem;
	uint8_t *highres_word;
	uint8_t *res_uv64;   // res_U_64 (512) then res_V_64 (512)
	uint8_t *sel1, *sel2;
	uint8_t *codebook1, *codebook2;
	uint32_t *words;     // `encode[]`
	void *pack_scratch;  // PackState (enc_pack.cuh)
	EncHdr *hdr;
};
