// enc_batch.cuh -- device workspace layout of a chunk of images being encoded.
// Planes live in per-kind arrays (slot = plane + zero guards, enc_img.cuh); all byte-sized
// per-image state lives in one "bytes" slot carved by the offsets below, each region followed
// by a guard gap that is never written (it is zeroed once at context creation).
#pragma once
#include "enc_img.cuh"

struct EncBatch {
	int16_t *y_proc, *y_jpeg, *y_aux, *y_ll1, *y_ll2s;
	int16_t *y_hq;           // q22/q23: kept first pass (256x512) + LL1 copy (256x256) + rebuilt LH1 (256x256), one plane slot
	int16_t *c_proc, *c_jpeg, *c_aux, *c_ll1, *c_ll2s;
	uint8_t *bytes;
	EncHdr *hdr;
};

#define ENC_GAP 512
#define ENC_ALIGN(x) (((x) + 15) & ~15)
#define ENC_NEXT(off, size) ENC_ALIGN((off) + (size) + ENC_GAP)

#define ENC_WORDS_CAP 131072                      // 32-bit words of prefix-code output per image
#define ENC_WORDS_BYTES (ENC_WORDS_CAP * 4)

enum : int {
	OFF_SCAN = 4096,
	OFF_TREE1 = ENC_NEXT(OFF_SCAN, NHW_SCAN_BYTES),
	OFF_CHRES = ENC_NEXT(OFF_TREE1, NHW_CAP_TREE1),
	OFF_LLCODE = ENC_NEXT(OFF_CHRES, 16384),
	OFF_EXW = ENC_NEXT(OFF_LLCODE, 49152),
	OFF_EXWUV = ENC_NEXT(OFF_EXW, 49152),
	OFF_RES1 = ENC_NEXT(OFF_EXWUV, 32768),
	OFF_RES1_BIT = ENC_NEXT(OFF_RES1, 65600),
	OFF_RES1_WORD = ENC_NEXT(OFF_RES1_BIT, 8224),
	OFF_RES3 = ENC_NEXT(OFF_RES1_WORD, 8224),
	OFF_RES3_BIT = ENC_NEXT(OFF_RES3, 65600),
	OFF_RES3_WORD = ENC_NEXT(OFF_RES3_BIT, 8224),
	OFF_RES4 = ENC_NEXT(OFF_RES3_WORD, 16448),
	OFF_RES5 = ENC_NEXT(OFF_RES4, 8192),
	OFF_RES5_BIT = ENC_NEXT(OFF_RES5, 65600),
	OFF_RES5_WORD = ENC_NEXT(OFF_RES5_BIT, 8224),
	OFF_RES6 = ENC_NEXT(OFF_RES5_WORD, 8224),
	OFF_RES6_BIT = ENC_NEXT(OFF_RES6, 65600),
	OFF_RES6_WORD = ENC_NEXT(OFF_RES6_BIT, 8224),
	OFF_CHARRES1 = ENC_NEXT(OFF_RES6_WORD, 8224),
	OFF_QSET3 = ENC_NEXT(OFF_CHARRES1, 512),
	OFF_TMP1 = ENC_NEXT(OFF_QSET3, 65536),
	OFF_TMP2 = ENC_NEXT(OFF_TMP1, 65600),
	OFF_TMP3 = ENC_NEXT(OFF_TMP2, 65600),
	OFF_HRMEM = ENC_NEXT(OFF_TMP3, 65600),
	OFF_HRWORD = ENC_NEXT(OFF_HRMEM, 32768),
	OFF_UV64 = ENC_NEXT(OFF_HRWORD, 16384),
	OFF_SEL1 = ENC_NEXT(OFF_UV64, 1024),
	OFF_SEL2 = ENC_NEXT(OFF_SEL1, 32800),
	OFF_BOOK1 = ENC_NEXT(OFF_SEL2, 32800),
	OFF_BOOK2 = ENC_NEXT(OFF_BOOK1, 1024),
	OFF_PACK = ENC_NEXT(OFF_BOOK2, 1024),
	OFF_WORDS = ENC_NEXT(OFF_PACK, 8192),
	ENC_BYTES_SLOT = ENC_NEXT(OFF_WORDS, ENC_WORDS_BYTES),
};
