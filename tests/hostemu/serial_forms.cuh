// tests/hostemu/serial_forms.cuh -- TEST TOOLING ONLY.
//
// First-generation, one-walk-per-image (or per row / per column) forms of encoder and decoder stages: straight restatements of the
// reference's loops that the CUDA kernels no longer call (they run the cell-group / row-mask / segment forms of
// nhwcodec_b200/csrc).  The host harness keeps them as the baseline its schedule checks compare those parallel forms with.
// Moved out of the product headers; grouped by the header they came from.
#pragma once

// ======== from nhwcodec_b200/csrc/dec_par.cuh ========

// D8: cell (r, j), 1 <= r, j <= 254, of the level-2 region of J (stride 512)
NHW_HD WfGeom dwf_shrink_geom() { return WfGeom{1, 254, 1, 254, 2}; }

// D11: pair p (columns 1+2p, 2+2p), 0 <= p <= 126, rows 1..254 of the reconstructed LL1 (stride 512)
NHW_HD WfGeom dwf_edge_geom() { return WfGeom{1, 254, 0, 127, 2}; }

// D16: chroma sharpen cell (r, j), 1 <= r, j <= 254 (stride 256)
NHW_HD WfGeom dwf_sharpen_geom() { return WfGeom{1, 254, 1, 254, 2}; }

// ======== from nhwcodec_b200/csrc/dec_stages.cuh ========

// ---- D2: inverse serpentine scan (nhw_decoder.c:71-91): coefficient stream -> transposed plane
NHW_HD void dec_y_descan_strip(const int16_t *coef, int16_t *J, int strip /* 0..127 */)
{
	const int16_t *s = coef + strip * 2048;
	int16_t *P = J + strip * 4;
	for (int k = 0; k < 256; k++) {
		int16_t *r0 = P + (2 * k) * YW, *r1 = r0 + YW;
		r0[0] = s[0]; r0[1] = s[1]; r0[2] = s[2]; r0[3] = s[3];
		r1[3] = s[4]; r1[2] = s[5]; r1[1] = s[6]; r1[0] = s[7];
		s += 8;
	}
}

NHW_HDN int dec_y_ll_image(const DecImg &im)
{
	for (int r = 0; r < 128; r++)
		for (int j = 0; j < 128; j++) im.jpeg[r * YW + j] = im.res_comp[r * 128 + j];
	return dec_y_ll_overrides(im);
}

// ---- D8: shrink isolated coefficients of the level-2 bands, in place (nhw_decoder.c:685-711)
// q <= 16 tolerates diagonal neighbours up to 16 (nhw_decoder.c:660-684)
NHW_HDN void dec_y_shrink_image(const DecImg &im)
{
	int16_t *J = im.jpeg;
	const int dg = im.d->quality <= 16 ? 16 : 8;
	for (int r = 1; r < 255; r++)
		for (int j = 1; j < 255; j++) {
			const int s = r * YW + j;
			if (nhw_iabs(J[s]) <= 8) continue;
			if (nhw_iabs(J[s - YW - 1]) > dg || nhw_iabs(J[s - YW]) > 8 || nhw_iabs(J[s - YW + 1]) > dg ||
			    nhw_iabs(J[s - 1]) > 8 || nhw_iabs(J[s + 1]) > 8 || nhw_iabs(J[s + YW - 1]) > dg ||
			    nhw_iabs(J[s + YW]) > 8 || nhw_iabs(J[s + YW + 1]) > dg)
				continue;
			if (r >= 128 || j >= 128) J[s] += J[s] > 0 ? -1 : 1;
		}
}

// ---- D10: residual add-backs on the reconstructed LL1 (nhw_decoder.c:721-787)
NHW_HDN void dec_y_addbacks_image(const DecImg &im)
{
	int16_t *P = im.proc;
	const int q = im.d->quality;
	auto at = [](uint16_t v) { return ((v & 65280) << 1) + (v & 255); };
	if (q >= 21) {
		for (int i = 0; i < im.list_len[2]; i++) P[at(im.list[2][i])] -= 3;
		for (int i = 0; i < im.list_len[3]; i++) P[at(im.list[3][i])] += 3;
	}
	if (q > 12) {
		const int e = q >= 18 ? 5 : q >= 15 ? 7 : 9;
		for (int i = 0; i < im.list_len[0]; i++) P[at(im.list[0][i])] -= e;
		for (int i = 0; i < im.list_len[1]; i++) P[at(im.list[1][i])] += e;
	}
	if (q >= 19) {
		for (int i = 0; i < im.list_len[5]; i++) { const int a = at(im.list[5][i]); P[a] -= 4; P[a + YW] -= 3; }
		for (int i = 0; i < im.list_len[4]; i++) { const int a = at(im.list[4][i]); P[a] += 4; P[a + YW] += 3; }
		for (int i = 0; i < im.list_len[6]; i++) { const int a = at(im.list[6][i]); P[a] += 2; P[a + YW] += 2; P[a + 2 * YW] += 2; }
		for (int i = 0; i < im.list_len[7]; i++) { const int a = at(im.list[7][i]); P[a] -= 2; P[a + YW] -= 2; P[a + 2 * YW] -= 2; }
	}
}

// ---- D11+D12: edge flags on LL1 (flagged cells carry +16000 while the pass runs, so later
// stencils see them), then the flag list in raster order (nhw_decoder.c:789-839)
NHW_HDN void dec_y_edge_flags_image(const DecImg &im)
{
	int16_t *P = im.proc;
	for (int r = 1; r < 255; r++)
		for (int j = 1; j < 254; j++) {
			int s = r * YW + j;
			const int res = dec_lap8(P, s, YW);
			j++; s++;
			const int cnt = dec_lap8(P, s, YW);
			if (res > 41 && res < 108 && cnt < 16) P[s - 1] += 16000;
			else if (res < -41 && res > -108 && cnt > -16) P[s - 1] += 16000;
			else if (cnt > 41 && cnt < 108 && res < 16) P[s] += 16000;
			else if (cnt < -41 && cnt > -108 && res > -16) P[s] += 16000;
		}
	int n = 0;
	for (int r = 1; r < 255; r++)
		for (int j = 0; j < 256; j++) {
			const int s = r * YW + j;
			if (P[s] > 10000) { im.flags[n++] = (uint16_t)((r << 8) + j); P[s] -= 16000; }
		}
	im.list_len[9] = n;
}

NHW_HDN void dec_y_smooth_flags_image(const DecImg &im) { dec_y_smooth_flags_plane(im, im.jpeg); }

// ---- chroma (nhw_decoder.c:895-1183 / 1185-1474), one component
NHW_HD void dec_c_descan_strip(const int16_t *coef, int16_t *J, int strip /* 0..31 */, int is_v)
{
	const int16_t *s = coef + is_v + strip * 4096;
	int16_t *P = J + strip * 8;
	for (int k = 0; k < 128; k++) {
		int16_t *r0 = P + (2 * k) * CW, *r1 = r0 + CW;
		for (int t = 0; t < 8; t++) r0[t] = s[2 * t];
		for (int t = 0; t < 8; t++) r1[7 - t] = s[16 + 2 * t];
		s += 32;
	}
}

// in-place 8-neighbour sharpen, raster order (nhw_decoder.c:1085-1109), then clip to 0..255
NHW_HDN void dec_c_sharpen_image(const DecImg &im)
{
	int16_t *P = im.cproc;
	const int thr = im.d->quality <= 14 ? 35 : 60;
	for (int r = 1; r < 255; r++)
		for (int j = 1; j < 255; j++) {
			const int s = r * CW + j;
			const int res = dec_lap8(P, s, CW);
			if (nhw_iabs(res) > thr) {
				if (res > 0) P[s] += res > 160 ? 3 : 2;
				else P[s] -= res < -160 ? 3 : 2;
			}
		}
	for (int i = 0; i < 65536; i++)
		if ((P[i] >> 8) != 0) P[i] = (int16_t)(P[i] < 0 ? 0 : 255);
}

// 2x upsample of the clipped 256x256 plane to 512x512 bytes: rows first, then columns, both
// (a+b+1)>>1 with the last row/column repeated (nhw_decoder.c:1137-1181)
NHW_HD void dec_c_upsample_row(const int16_t *P, uint8_t *out, int y /* 0..511 */)
{
	const int r = y >> 1;
	auto v = [&](int c) -> int {
		if ((y & 1) == 0 || r == 255) return P[r * CW + c];
		return (P[r * CW + c] + P[(r + 1) * CW + c] + 1) >> 1;
	};
	uint8_t *o = out + y * 512;
	for (int c = 0; c < 255; c++) {
		const int a = (uint8_t)v(c), b = (uint8_t)v(c + 1);
		o[2 * c] = (uint8_t)a;
		o[2 * c + 1] = (uint8_t)((a + b + 1) >> 1);
	}
	o[510] = o[511] = (uint8_t)v(255);
}

// ======== from nhwcodec_b200/csrc/enc_c.cuh ========

// Row form of offsetUV.  The only cross-row access is the un-guarded look at P[i+1] from the last
// column (image_processing.c:151-154): `next0` is the next row's first cell BEFORE it is quantised
// (0 after the last row).
NHW_HD void c_offset_quant_row(const EncImg &im, int m2, int r, int next0)
{
	int16_t *P = im.cproc + r * CW;
	for (int c = 0; c < 256; c++) {
		const bool inrow = c < 255;
		int a = P[c];
		if (a > 10000) {
			int b = a == 12400 ? 124 : a == 12600 ? 126 : a == 12900 ? 122 : a == 13000 ? 130 : -1;
			if (b >= 0) { P[c] = (int16_t)b; continue; }
		}
		if (a > 127) {
			int k = ((a & 0xfff8) - 128) >> 3;
			P[c] = NHW_EXTRA1(k > 18 ? 18 : k);
			continue;
		} else if (a < -127) {
			int k = (((-a) & 0xfff8) - 128) >> 3;
			P[c] = NHW_EXTRA2(k > 18 ? 18 : k);
			continue;
		}
		const int nxt = inrow ? (int)P[c + 1] : next0;
		const bool neg = a < 0;
		if (a == -7 || a == -8) {
			if (inrow && (nxt == -7 || nxt == -8)) { P[c] = 120; P[c + 1] = 120; c++; continue; }
		}
		if (neg) {
			a = -a;
			if (nxt < 0 && nxt > -8) { if ((a & 7) < 6) a &= 504; }
			else if ((a & 7) < 7) a &= 504;
			a = -a;
		} else if (a > 6 && (a & 7) >= 6) {
			if (inrow && nxt == 7) P[c + 1] = 8;
		}
		if (a < m2 && a > -m2) { P[c] = 128; continue; }
		P[c] = (int16_t)((a + 128) & 248);
	}
}

// ---- chroma scan: 8-column strips, two rows per step, U on even / V on odd bytes from 262144
NHW_HD void c_scan_strip(const EncImg &im, int strip /* 0..31 */, int is_v)
{
	const int16_t *P = im.cproc + strip * 8;
	uint8_t *s = im.scan + 262144 + is_v + strip * 4096;
	for (int k = 0; k < 128; k++) {
		const int16_t *r0 = P + (2 * k) * CW, *r1 = r0 + CW;
		for (int t = 0; t < 8; t++) s[2 * t] = (uint8_t)r0[t];
		for (int t = 0; t < 8; t++) s[16 + 2 * t] = (uint8_t)r1[7 - t];
		s += 32;
	}
}

// ======== from nhwcodec_b200/csrc/enc_hq.cuh ========

NHW_HDN void hq_e17_image(const EncImg &im)
{
	for (int r = 0; r < 256; r++)
		for (int j = 0; j < 254; j++) {
			int d[3];
			if (!hq_e17_delta(im.ll1[r * 256 + j], d)) continue;
			for (int k = 0; k < 3; k++) im.hq_fo[j * 256 + r + k] = (int16_t)(im.hq_fo[j * 256 + r + k] + d[k]);
		}
}

NHW_HD void hq_tag_row(const EncImg &im, int q, int r)
{
	for (int t = 0; t < 256; t++) hq_tag_pair(im, q, r, t);
}

// ---- everything after the tags, serially (the lists are short): char_res1, high_qsetting3, and the
// res6 position list through the same pruning / packing as res1 (y_e18_finish_list_image, which = 6)
NHW_HDN int hq_lists_image(const EncImg &im, int q)
{
	EncHdr *h = im.hdr;
	int total = 0;
	for (int r = 0; r < 256; r++) { int nw; total += hq_collect_row(im, r, nullptr, nullptr, nw); }
	if (total + 16 > NHW_CAP_LIST) return NHW_ERR_OVERFLOW_DEV;
	int count = 0, e = 0, res = 0;
	for (int r = 0; r < 256; r++) {
		int nw;
		count += hq_collect_row(im, r, im.tmp1 + count, im.tmp3 + e, nw);
		e += nw;
		const uint8_t *tag = im.hq_tag + r * 512;
		for (int k = 0; k < 2; k++) {
			const int g = tag[254 + k];
			if (g == 1 || g == 2) {
				if (res >= NHW_CAP_CHAR_RES1) return NHW_ERR_OVERFLOW_DEV;
				im.char_res1[res++] = (uint16_t)(r * 256 + 2 * k + (g - 1));
			}
		}
	}
	h->char_res1_len = res;
	int n3 = 0;
	if (q > 22)
		for (int i = 0; i < 131072; i++) {
			const int g = im.hq_tag[i];
			if (g == 3 || g == 4) {
				if (n3 >= NHW_CAP_QSETTING3) return NHW_ERR_OVERFLOW_DEV;
				im.qsetting3[n3++] = (uint32_t)(i << 1) + (g == 4 ? 1u : 0u);
			}
		}
	h->qsetting3_len = n3;
	y_e18_finish_list_image(im, 6, count, e);
	return 0;
}

// ======== from nhwcodec_b200/csrc/enc_ll_par.cuh ========

// byte stores of one non-escape cell
NHW_HD void ll2_bytes_store(const EncImg &im, int a, int scan)
{
	if (scan > 255) scan = 255;
	else if (scan < 0) scan = 0;
	im.ch_res[a] = (uint8_t)scan;
	im.tree1[a] = (uint8_t)(scan & 254);
}

// escape cell a (visited in raster order): copies the previous byte, appends to exw_Y.  e = list length
NHW_HD void ll2_bytes_escape(const EncImg &im, int a, int scan, int &e)
{
	im.exw[e++] = (uint8_t)(a >> 7);
	if (scan > 255) {
		im.exw[e++] = (uint8_t)((a & 127) + 128);
		const int y = scan - 255;
		im.exw[e++] = (uint8_t)(y > 255 ? 255 : y);
	} else {
		im.exw[e++] = (uint8_t)(a & 127);
		im.exw[e++] = (uint8_t)(scan < -255 ? 255 : -scan);
	}
	im.tree1[a] = im.tree1[a - 1];
	im.ch_res[a] = im.tree1[a - 1];
}

// serial reference of the whole coder built on the step function (host harness / fallback shape)
NHW_HDN int ll_dpcm_luma_steps(const EncImg &im, const uint8_t *x, int q)
{
	EncHdr *h = im.hdr;
	const int N = 16384;
	int a8 = 0, y16 = 0;
	for (int i = 1; i < N; i++)
		if (x[i] == x[i - 1] && (i == 1 || x[i - 1] != x[i - 2])) ll_stats_run(x, i, N, a8, y16);
	const int mode = y16 > 299 ? 2 : (a8 + y16 > 179 ? 1 : 0);
	uint8_t *out = im.llcode;
	out[0] = x[0];
	int o = 1, nmem = 0;
	for (int i = 1; i < N;) {
		const LlStep s = ll_dpcm_step(x, i, mode, q);
		out[o++] = s.b[0];
		if (s.nbytes == 2) out[o++] = s.b[1];
		if (s.raw) { im.highres_word[nmem] = im.ch_res[i]; im.highres_mem[nmem++] = (uint16_t)i; }
		i = s.next;
	}
	h->highres_comp_len = nmem;
	h->highres_mem_len = nmem;
	h->res_low = mode;
	h->y_res_comp = o;
	return mode;
}

// ======== from nhwcodec_b200/csrc/enc_par.cuh ========

// offsetY_recons256 pattern substitution: region A then region B (image_processing.c:2759-2849).
// cell (r,j) reads (r,j-1..j+1),(r+1,j-1..j); writes (r,j-1),(r,j),(r+1,j-1),(r+1,j): skew 3.
NHW_HD WfGeom wf_recons_patterns_geom(int region) { return region == 0 ? WfGeom{0, 128, 129, 126, 3} : WfGeom{128, 127, 1, 254, 3}; }

// returns the number of columns consumed (1 or 2)
NHW_HD int wf_recons_patterns_cell(const EncImg &im, int r, int j)
{
	int a = r * YW + j, jj = j;
	recons_pattern_cell(im.proc, im.jpeg, a, jj);
	return jj - j + 1;
}

// offsetY_recons256 isolated-coefficient shrink (image_processing.c:3162-3187): reads all 8
// neighbours, row above edited, row below not; writes (r,j): skew 2.
NHW_HD WfGeom wf_shrink_geom() { return WfGeom{1, 254, 1, 254, 2}; }

// clean-up of the level-1 detail bands (nhw_encoder.c:1923-2098), three passes.
// cell (r,j) reads (r-1,j) edited, (r+1,j) un-edited, (r,j-1..j+2); writes (r,j),(r,j+1): skew 2.
NHW_HD WfGeom wf_e20_geom(int pass) { return pass == 0 ? WfGeom{1, 254, 257, 254, 2} : pass == 1 ? WfGeom{256, 255, 1, 255, 2} : WfGeom{256, 255, 257, 254, 2}; }

NHW_HD int wf_e20_cell(const EncImg &im, int q, int ratio, int pass, int r, int j)
{
	int yw, yw2, lo, jmax;
	if (pass == 0) { if (q > 22) { yw = 8; yw2 = 4; } else { yw = 9; yw2 = 9; } lo = ratio - 2; jmax = 510; }
	else if (pass == 1) { if (q > 22) { yw = 8; yw2 = 4; } else if (q > 17) { yw = 8; yw2 = 9; } else { yw = 9; yw2 = 9; } lo = ratio - 2; jmax = 254; }
	else { yw = q > 22 ? 8 : 11; yw2 = yw; lo = ratio - 1; jmax = 510; }
	e20_cell(im.proc, r * YW + j, j, jmax, lo, yw, yw2, pass);
	return 1;
}

// offsetY pattern marks in the level-2 region (image_processing.c:239-290): same footprint as
// the recons patterns: skew 3.
NHW_HD WfGeom wf_offset_patterns_geom() { return WfGeom{0, 256, 1, 254, 3}; }

// offsetY loop 4 (image_processing.c:312-519), one row.  `next0` is the NOT YET QUANTISED first
// cell of the next row (0 after the last row): the one place the reference looks across the
// row end without a bounds test.
NHW_HD void y_offset_quant_row(const EncImg &im, int m1, int r, int next0)
{
	int16_t *P = im.proc + r * YW;
	for (int c = 0; c < 512; c++) {
		const bool inrow = c < 511;
		const int nxt = inrow ? (int)P[c + 1] : next0;
		int a = P[c];
		if (a > 10000) {
			int b = a == 10100 ? 128 : a == 12700 ? 127 : a == 12900 ? 129 : a == 10204 ? 125 : a == 10300 ? 126 :
			        a == 12100 ? 121 : a == 12200 ? 122 : -1;
			if (b >= 0) { P[c] = (int16_t)b; continue; }
		}
		if (a > 127) {
			int k = ((a & 0xfff8) - 128) >> 3;
			P[c] = NHW_EXTRA1(k > 18 ? 18 : k);
			continue;
		} else if (a < -127) {
			int k = (((-a) & 0xfff8) - 128) >> 3;
			P[c] = NHW_EXTRA2(k > 18 ? 18 : k);
			continue;
		}
		if (a < -12 && ((-a) & 7) == 6) {
			if (inrow && nxt == -7) P[c + 1] = -9;
		}
		if (a < 0) {
			const int nx = inrow ? (int)P[c + 1] : next0;   // may just have become -9
			if (a == -7 && nx == 8 && inrow) { P[c] = -8; a = -8; }
			a = -a;
			if (a > 14 && (a & 7) == 7 && nx > 0 && nx < 8) a -= 2;
			if ((a & 7) < 7) a &= 504;
			a = -a;
		} else if (a == 8 && nxt == -7 && inrow) P[c + 1] = -8;
		else if (a > 12 && (a & 7) >= 6) {
			if (inrow && nxt == 7) P[c + 1] = 9;
		}
		if (a < m1 && a > -m1) { P[c] = 128; continue; }
		P[c] = (int16_t)((a + 128) & 248);
	}
}

// ---- LL2 part of offsetY_recons256 (image_processing.c:2610-2737) in parallel form --------------
// P = LL2 band at row stride PS (shared-memory copy), J = im_jpeg (stride 512).
//   1. quad tagging: rows independent
//   2. main loop: cell (r,j) reads (r,j..j+2),(r+1,j..j+2),(r+2,j),(r+3,j) and writes (r,j),(r,j+1),
//      (r+1,j): the row above must be 3 columns ahead -> wavefront, skew 3
//   3. second call only: un-tag + copy to jpeg (cells independent), then the highres_mem fix-ups
NHW_HD WfGeom wf_ll2_geom() { return WfGeom{0, 128, 0, 128, 3}; }

NHW_HD void y_recons_ll2_tail_row(int16_t *P, int PS, int16_t *J, int16_t *tmp, int r)
{
	int a = r * PS, aj = r * YW, t = r * 128;
	for (int j = 0; j < 128; j++, a++, aj++, t++) {
		if (P[a] < 10000) {
			tmp[t] = P[a];
			J[aj] = (P[a] >= 0 && P[a] < 256) ? (int16_t)(P[a] & 65534) : P[a];
		} else {
			P[a] -= 16000;
			tmp[t] = P[a];
			J[aj] = P[a];
		}
	}
}

// ======== from nhwcodec_b200/csrc/enc_point.cuh ========

// position of chroma cell (row, col) of plane is_v in the scan buffer: 8-column strips, U on even bytes
NHW_HD int c_scan_pos(int row, int col, int is_v)
{
	const int t = col & 7;
	return 262144 + is_v + (col >> 3) * 4096 + (row >> 1) * 32 + ((row & 1) ? 16 + 2 * (7 - t) : 2 * t);
}

// ======== from nhwcodec_b200/csrc/enc_y1.cuh ========

NHW_HDN void y_recons_ll2_image(const EncImg &im, int q, int part)
{
	y_recons_ll2_core(im.proc, YW, im.jpeg, im.aux, im.highres_mem, im.hdr->highres_mem_len, q, part);
}

// ---- offsetY_recons256, second call only: shrink isolated reconstructed coefficients,
// in place and in raster order (image_processing.c:3162-3187, q>16 branch)
NHW_HDN void y_recons_shrink_image(const EncImg &im, int q = 20)
{
	int16_t *J = im.jpeg;
	const int dg = q <= 16 ? 16 : 8;   // q <= 16: diagonal neighbours only count from 16 up (image_processing.c:3137-3160)
	for (int r = 1; r < 255; r++) {
		int e = r * YW + 1;
		for (int j = 1; j < 255; j++, e++) {
			if (nhw_iabs(J[e]) < 8) continue;
			if (nhw_iabs(J[e - YW - 1]) >= dg || nhw_iabs(J[e - YW]) >= 8 || nhw_iabs(J[e - YW + 1]) >= dg ||
			    nhw_iabs(J[e - 1]) >= 8 || nhw_iabs(J[e + 1]) >= 8 || nhw_iabs(J[e + YW - 1]) >= dg ||
			    nhw_iabs(J[e + YW]) >= 8 || nhw_iabs(J[e + YW + 1]) >= dg)
				continue;
			if (r >= 128 || j >= 128) J[e] += J[e] > 0 ? -1 : 1;
		}
	}
}

NHW_HD void y_e6d_correct_row(const EncImg &im, int r)
{
	y_e6d_correct_cells(im.proc + r * YW, im.jpeg + r * YW, im.ll1 + r * 256);
}

// ======== from nhwcodec_b200/csrc/enc_y2.cuh ========

NHW_HDN void y_e16_residual_col_w(const EncImg &im, int q, int j, const int16_t *Pn, const int16_t *Ln, const uint8_t *lut = nullptr)
{
	y_e16_residual_col_t<false>(im, q, j, Pn, Ln, lut, 0);
}

NHW_HDN void y_e16b_classify_image(const EncImg &im, int q)
{
	int w1 = 0, w3 = 0, w5 = 0;
	for (int j = 0; j < 256; j++) y_e16b_classify_col_w(im, q, j, w1, w3, w5);
	im.hdr->res1_word_len = w1;
	im.hdr->res3_word_len = w3;
	im.hdr->res5_word_len = w5;
}

NHW_HDN void y_e18_pack_list_image(const EncImg &im, int which)
{
	uint8_t *pos = im.tmp1, *wrd = im.tmp3;
	int count = 0, e = 0;
	for (int row = 0; row < 256; row++) {
		const int n = y_e18_collect_row(im, which, row, pos + count, wrd + e);
		count += n + 1;
		e += n;
	}
	y_e18_finish_list_image(im, which, count, e);
}

NHW_HDN void y_e20_cleanup_image(const EncImg &im, int q, int ratio)
{
	int16_t *P = im.proc;
	int yw, yw2;
	if (q > 22) { yw = 8; yw2 = 4; } else { yw = 9; yw2 = 9; }
	for (int r = 1; r < 255; r++)
		for (int j = 257; j < 511; j++) e20_cell(P, r * YW + j, j, 510, ratio - 2, yw, yw2, 0);
	if (q > 22) { yw = 8; yw2 = 4; } else if (q > 17) { yw = 8; yw2 = 9; } else { yw = 9; yw2 = 9; }
	for (int r = 256; r < 511; r++)
		for (int j = 1; j < 256; j++) e20_cell(P, r * YW + j, j, 254, ratio - 2, yw, yw2, 1);
	yw = q > 22 ? 8 : 11;
	for (int r = 256; r < 511; r++)
		for (int j = 257; j < 511; j++) e20_cell(P, r * YW + j, j, 510, ratio - 1, yw, yw, 2);
}

// ======== from nhwcodec_b200/csrc/enc_y3.cuh ========

// ---- offsetY loop 1 (image_processing.c:194-237): neighbouring multiples of 8 in the detail
// bands; flat raster order (col 0 looks at the previous row's last, already-visited cell).
NHW_HDN void y_offset_pairs_image(const EncImg &im)
{
	int16_t *P = im.proc;
	for (int i = 0; i < 4 * 65536; i++) {
		const int col = i & 511;
		if (!(i >= 2 * 65536 || col >= 256)) continue;
		if (!(P[i] > 7 && P[i + 1] > 7 && col < 511)) continue;
		int a = P[i];
		if ((a & 7) || (P[i + 1] & 7)) continue;
		if (a > 15) {
			if (i > 0) {
				if (P[i - 1] <= 0) P[i]--;
				else if (P[i + 1] > 15) {
					if (col < 510 && P[i + 2] <= 0) P[i + 1]--;
				}
			}
		} else if (P[i + 1] > 15) {
			if (col < 510 && P[i + 2] <= 0) P[i + 1]--;
		}
	}
}

// ---- E24: peephole passes over the 262144 luma bytes
NHW_HDN void y_peephole_image(const EncImg &im)
{
	uint8_t *s = im.scan;
	const int N = 4 * 65536;
	for (int i = 0; i < N - 4; i++) {
		if (s[i] != 128 && s[i + 1] == 128) {
			if (s[i + 2] == 128) {
				if (s[i + 3] == 128) {
					int x = s[i], y = s[i + 4];
					if ((x == 136 || x == 120) && (y == 136 || y == 120)) {
						s[i] = (uint8_t)(132 + (x == 120 ? 2 : 0) + (y == 120 ? 1 : 0));
						s[i + 4] = 201;
						i += 4;
					} else i += 3;
				} else i += 2;
			} else i++;
		}
	}
	s[0] = s[1] = s[2] = s[3] = 128;
	s[N - 4] = s[N - 3] = s[N - 2] = s[N - 1] = 128;
	int sel1 = 0, sel2 = 0;
	for (int i = 4; i < N - 4; i++) {
		if (s[i] != 136 && s[i] != 120) continue;
		const bool nxt = (s[i + 1] == 120 || s[i + 1] == 136);
		if (s[i + 2] == 128 && nxt && s[i - 1] == 128 && s[i - 2] == 128 && s[i - 3] == 128 && s[i - 4] == 128) {
			s[i + 1] = (uint8_t)(s[i + 1] == 120 ? 157 : 159);
			sel2++;
		} else if (s[i - 1] == 128 && nxt && s[i + 2] == 128 && s[i + 3] == 128 && s[i + 4] == 128 && s[i + 5] == 128) {
			s[i + 1] = (uint8_t)(s[i + 1] == 120 ? 157 : 159);
			sel2++;
		} else if (s[i - 1] == 128 && s[i - 2] == 128 && s[i - 3] == 128 && s[i - 4] == 128 && s[i + 1] == 128) {
			s[i] = (uint8_t)(s[i] == 136 ? 153 : 155);
			sel1++;
		} else if (s[i - 1] == 128 && s[i + 1] == 128 && s[i + 2] == 128 && s[i + 3] == 128 && s[i + 4] == 128) {
			s[i] = (uint8_t)(s[i] == 136 ? 153 : 155);
			sel1++;
		}
	}
	im.hdr->select1 = sel1;
	im.hdr->select2 = sel2;
	for (int i = 0, count = 0; i < N; i++) {
		while (s[i] == 128 && s[i + 1] == 128) {
			count++;
			if (count > 255) {
				for (int k = 0; k < 4; k++) {
					if (s[i + k] == 153) s[i + k] = 124;
					else if (s[i + k] == 155) s[i + k] = 123;
				}
				i--;
				count = 0;
			} else i++;
		}
		if (count >= 252) {
			if (s[i + 1] == 153) s[i + 1] = 124;
			else if (s[i + 1] == 155) s[i + 1] = 123;
		}
		count = 0;
	}
}

// ======== from nhwcodec_b200/csrc/pre_lowq.cuh ========

// ---- the whole stage ---------------------------------------------------------------------------------------------
// Y: in/out.  O, K: scratch planes (O receives the copy).  M: 262144 bytes of scratch.
NHW_HDN void pre_low_image(int16_t *Y, int16_t *O, int16_t *K, uint8_t *M, int q)
{
	const PreLowParams p = pre_low_params(q);
	for (int i = 0; i < PW * PW; i++) { O[i] = Y[i]; M[i] = 0; }
	// the kernel plane's border is never written by walk A and is read by the later walks: it reads as zero
	for (int i = 0; i < PW; i++) { K[i] = 0; K[511 * PW + i] = 0; K[i * PW] = 0; K[i * PW + 511] = 0; }
	pre_low_walk_a(O, K, p);
	pre_low_walk_b(Y, O, K, M, p);
	pre_low_walk_c(Y, K, M, p);
	for (int r = 1; r < 511; r++) pre_low_walk_d_row(Y, K, M, p, r);
}
