// pre_lowq.cuh -- luma pre-sharpening at q <= 16 (pre_processing, encoder/image_processing.c:558-2426 with
// quality_setting <= LOW4).
//
// At these settings the reference's pre-sharpening is a raster-ordered state machine: four walks over the
// 510 x 510 interior, the first three of which carry state (counters that throttle how often a rule may fire)
// from one pixel pair to the next through the whole image.  There is nothing to parallelise inside an image
// beyond the Laplacian itself; the batch is the parallel axis (one walker per image).
//
//   walk A  (:601-764)   kernel value of every pixel: |Laplacian| * 15 + sum of absolute differences, divided by 16
//                        with the remainder carried forward; values that cross the sharpness threshold only through
//                        the carried remainder are replaced by markers (+-20000, 7000) on a throttle
//   walk B  (:770-1992)  per horizontal pair: optional 5-tap smoothing (q <= 14), the throttled +-1 / +-2 nudges
//                        (PairThrottle below), opposite-sign pairs, and the mid-range rule of the q > 16 path
//                        (q15, q16 and q8..q10 only)
//   walk C  (:1994-2274) resolves the markers, then a strong/weak pair rule that also reaches into the row above;
//                        the column cursor backs up and re-visits pixels (per-row cursor state)
//   walk D  (:2276-2420) remaining like-signed pairs, skipping pixels walks B and C already touched
//
// Planes: Y = the luma plane being sharpened (in/out), O = a copy of it taken before the stage, K = kernel values,
// M = per-pixel "touched" marks (0 none, 1 walk C, 2 raised by walk B, 3 lowered by walk B).  All 512 x 512, stride 512.
#pragma once
#include "pre_core.cuh"

#define PW 512

struct PreLowParams { int sharp, sharp2, smooth_below, smooth_on, midrange_on; };
NHW_HD PreLowParams pre_low_params(int q)
{
	// sharpness per quality (image_processing.c:573-588); the smoothing window bound n1 (:592-599)
	const int sharp_of[17] = {0, 48, 45, 36, 24, 24, 0, 0, 0, 1, 17, 35, 41, 44, 49, 54, 59};
	const int n1_of[17] = {36, 60, 56, 36, 36, 36, 36, 6, 10, 24, 36, 36, 36, 36, 36, 36, 36};
	PreLowParams p;
	const int k = q < 1 ? 1 : q > 16 ? 16 : q;
	p.sharp = sharp_of[k];
	p.sharp2 = p.sharp < 10 ? 10 : p.sharp;
	p.smooth_below = n1_of[k];
	p.smooth_on = q <= 14;
	p.midrange_on = q > 14 || (q <= 10 && q > 7);
	return p;
}

// ---- walk A --------------------------------------------------------------------------------------------------
struct PreWalkA {
	int carry;                 // remainder of the /16, 0..15
	int neg_phase, pos_phase;  // 0: next crossing becomes a marker; 1, 2: it keeps its value
	int neg_skip, pos_skip;    // slow counters that stretch the phases (cycle 0,1,2,3)
	int alt, alt2;             // toggles driven by what the left neighbour turned into
	int first21;               // occurrences of the value sharp2 + 21
	int edge_bumps;            // the first three values equal to -sharp2 are pushed to -sharp2 - 1
};

NHW_HD int pre_lap(const int16_t *O, int s, int &sad)
{
	const int c = O[s];
	const int d0 = c - O[s - 1], d1 = c - O[s + 1], d2 = c - O[s - PW], d3 = c - O[s + PW];
	const int d4 = c - O[s - PW + 1], d5 = c - O[s - PW - 1], d6 = c - O[s + PW - 1], d7 = c - O[s + PW + 1];
	sad = nhw_iabs(d0) + nhw_iabs(d1) + nhw_iabs(d2) + nhw_iabs(d3) + nhw_iabs(d4) + nhw_iabs(d5) + nhw_iabs(d6) + nhw_iabs(d7);
	return d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
}

// Walk A in two parts.  (1) The plain kernel values: sign(lap) * ((15 |lap| + sad + carry) >> 4) with the carried
// remainder -- the q > 16 recurrence, which the CUDA path evaluates as a scan (front.cu: k_pre_energy / k_pre_chain /
// k_pre_apply); this serial form is the host harness's.  (2) The marker rules: they only look at a pixel whose plain
// value satisfies pre_low_is_event, at its Laplacian, and at the left neighbour's CURRENT kernel value, and they never
// touch the remainder chain -- so they run as a walk over those pixels only, in raster order (pre_low_event).
NHW_HDN void pre_low_kernel_plane(const int16_t *O, int16_t *K)
{
	int carry = 0;
	for (int r = 1; r < 511; r++)
		for (int j = 1, s = r * PW + 1; j < 511; j++, s++) {
			int sad;
			const int lap = pre_lap(O, s, sad);
			if (lap == 0) { K[s] = 0; carry = 0; continue; }
			const int acc = 15 * nhw_iabs(lap) + sad + ((carry + 2) >> 2);
			carry = acc & 15;
			K[s] = (int16_t)(lap < 0 ? -(acc >> 4) : (acc >> 4));
		}
}

NHW_HD bool pre_low_is_event(int kv, int s2)
{
	const int k = nhw_iabs(kv);
	return (k > s2 && k <= s2 + 20) || kv == s2 + 21 || kv == -s2;
}

// pixel s (column j) holds the plain kernel value K[s] and pre_low_is_event says yes
NHW_HD void pre_low_event(PreWalkA &w, const PreLowParams &p, const int16_t *O, int16_t *K, int s, int j)
{
	const int s2 = p.sharp2, half = p.sharp >> 1;
	int sad;
	const int lap = pre_lap(O, s, sad);
	const int mag = nhw_iabs(lap);
	int k = nhw_iabs(K[s]);
	const int left = j > 1 ? K[s - 1] : 0;
	if (lap < 0) {
		if (k == s2 && w.edge_bumps < 3) { k = s2 + 1; w.edge_bumps++; }
		int out = -k;
		if (mag <= s2 && k > s2 && k <= s2 + 20) {   // crossed the threshold through the carried remainder only
			if (j > 1 && nhw_iabs(left) <= half) w.neg_phase = 0;
			if (w.neg_phase == 0) { out = -20000; w.neg_phase = 1; }
			else if (w.neg_skip == 0) { w.neg_phase = 0; w.neg_skip = 1; }
			else if (w.neg_phase == 1) w.neg_phase = 2;
			else { w.neg_phase = 0; w.neg_skip = w.neg_skip == 3 ? 0 : w.neg_skip + 1; }
		}
		K[s] = (int16_t)out;
	} else if (lap > 0) {
		int out = k;
		if (mag <= s2 && k > s2 && k <= s2 + 20) {
			auto toggle = [&]() {
				if (!w.alt) { w.pos_phase = 0; if (!w.pos_skip) w.pos_skip = 1; w.alt = 1; }
				else w.alt = 0;
			};
			if (j > 1) {
				if (nhw_iabs(left) <= half) w.pos_phase = 0;
				else if (nhw_iabs(left) > 10000 || left == s2 + 21) toggle();
				else if (left == -(s2 + 21)) {
					if (!w.alt2) w.alt2 = 1;
					else { toggle(); w.alt2 = w.alt2 == 1 ? 2 : 0; }
				} else if (left == s2 + 22) K[s - 1] = 7000;
			}
			if (w.pos_phase == 0) { out = 20000; w.pos_phase = 1; }
			else if (w.pos_skip == 0) { w.pos_phase = 0; w.pos_skip = 1; }
			else if (w.pos_phase == 1) w.pos_phase = 2;
			else { w.pos_phase = 0; w.pos_skip = w.pos_skip == 3 ? 0 : w.pos_skip + 1; }
		} else if (k == s2 + 21) {
			if (w.first21 == 0) out = 7000;
			w.first21++;
		}
		K[s] = (int16_t)out;
	}
}

NHW_HDN void pre_low_walk_a(const int16_t *O, int16_t *K, const PreLowParams &p)
{
	PreWalkA w = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	pre_low_kernel_plane(O, K);
	for (int r = 1; r < 511; r++)
		for (int j = 1, s = r * PW + 1; j < 511; j++, s++)
			if (pre_low_is_event(K[s], p.sharp2)) pre_low_event(w, p, O, K, s, j);
}

// ---- walk B: the throttle ---------------------------------------------------------------------------------------
// A pair alternates between a "strong" turn (+-2 nudges, kernel values consumed) and "weak" turns (+-1); how many
// weak turns follow a strong one, and which variant of the strong rule applies, is steered by the counters below.
// The names follow their role where one could be made out; the rest keep the reference's numbering (n[k] = tk,
// u[k] = wk) because their only meaning is the schedule they produce.
struct PairThrottle {
	int n[45];
	int u[9];
	NHW_HD void init()
	{
		for (int k = 0; k < 45; k++) n[k] = 0;
		for (int k = 0; k < 9; k++) u[k] = 0;
		n[6] = 8; n[10] = 10; n[11] = 15; n[18] = 8; n[44] = 2; u[3] = 20;
	}
};

// the nine-state sequencer that re-arms the weak-turn budget once n[7] has reached 4 (image_processing.c:1197-1466)
NHW_HD void throttle_sequencer(PairThrottle &t)
{
	int *n = t.n, *u = t.u;
	const int st = n[16];
	// budget pair (n10, n11) per state: states 0, 2, 4, 5 give (10, 15), the others (8, 12)
	const bool wide = st == 0 || st == 2 || st == 4 || st == 5;
	n[10] = wide ? 10 : 8;
	n[11] = wide ? 15 : 12;
	if (st == 0) {
		n[16] = 1;
		if ((u[7] == 2 || u[7] == 4) && n[24] == 14) { if (u[7] == 2) n[1] = 2000005; }
		else { n[4] = 1000000; n[1] = 9; }
	} else if (st == 1) {
		n[16] = 2;
		u[5]++;
		if (u[5] == 3 && n[1] > 0 && n[1] < 30) n[1] = (-n[1]) >> 2;
		else { n[4] = 10; n[1] += 2; }
	} else if (st == 2) {
		n[16] = 3; n[4] = 1000000;
		u[6]++;
		if (u[6] == 6 || u[6] == 10) n[1] = 10;
	} else if (st == 3) { n[16] = 4; n[4] = 8; n[1] -= 4; }
	else if (st == 4) n[16] = 5;
	else if (st == 5) { n[16] = 6; n[4] = 10; n[1] = 2000000; }
	else if (st == 6) { n[16] = 7; n[4] = 8; n[1] = 3000000; }
	else if (st == 7) { n[16] = 8; n[4] = 1000000; }
	else if (st == 8) {
		// sixteen sub-steps; next state, and what happens to n4 / n1 on the way
		const int sub = n[24];
		const int next16[16] = {1, 2, 1, 2, 1, 0, 3, 3, 1, 8, 1, 0, 1, 0, 1, 0};
		n[16] = next16[sub & 15];
		n[24] = sub == 15 ? 12 : sub + 1;
		if (sub == 0 || sub == 2 || sub == 9) n[4] = 1000000;
		else if (sub == 4) n[1] = 2999998;
		else if (sub == 7) n[1] = 7;
		else if (sub == 10) { n[4] = 8; n[1] = 11; }
		else if (sub == 14) { u[7]++; n[1] = u[2] == 0 ? 1999978 : u[2] == 1 ? 1999982 : 1999993; }
		else if (sub == 15) { n[1] = (u[2] == 1 || u[2] == 3) ? -5 : 2000005; u[2]++; }
	}
}

// budget exhausted (or forced): start the next cycle (image_processing.c:1063-1472)
NHW_HD void throttle_rearm(PairThrottle &t)
{
	int *n = t.n, *u = t.u;
	if (!n[6]) {
		n[6] = 1; n[14] = 0;
		if (!n[22]) n[7]++;
		if (n[22] == 1) n[22] = 0;
	} else {
		n[6]++; n[1]++;
		if (n[4] > 900000 && n[1] == 12) n[4] = 8;
		if (n[1] > 3000000) { n[1] = 12; n[4] = 8; }
		else if (n[1] > 2000006 && n[1] < 2500000) { n[1] = 14; n[4] = 10; }
		if (!n[15]) { n[14] = 1; n[15] = 1; }
		else { n[14] = 0; n[15]++; if (n[15] > 9) n[15] = 0; }
		if (n[6] > 15 && n[7] < 4) { n[6] = 0; if (n[19] > 0) n[20]++; }
	}
	if (n[4] == 8 || (n[4] == 10 && u[3] > 16)) {
		if (u[3] < 21) { n[4] = 0; u[3]++; }
		else if (n[4] == 8) u[3] = 0;
		else if (u[4] < 2) { n[4] = 8; n[1] = 12; u[4]++; }
		else { n[4] = 0; u[4] = 0; }
	} else n[4] = 0;
	n[8] = 0; n[5] = 0; n[12] = 0;
	if (n[7] == 3) {
		if (!n[6]) { n[10] = 10; n[11] = 15; } else { n[10] = 8; n[11] = 12; }
	} else if (n[7] == 1) {
		if (n[9] < 2) { n[10] = 10; n[11] = 15; n[9]++; }
		else { n[10] = 8; n[11] = 12; n[9]++; if (n[9] >= 3) n[9] = 0; }
	} else if (n[7] == 2) { n[10] = 8; n[11] = 12; }
	else if ((n[6] == 10 || n[6] == 11) && !n[7]) { n[10] = 6; n[11] = 9; }
	else if (n[7] >= 4) throttle_sequencer(t);
	else { n[10] = n[10] == 8 ? 10 : 8; n[11] = n[11] == 12 ? 15 : 12; }
}

// a weak turn with the budget counter n1 at 15 or more (image_processing.c:1473-1516)
NHW_HD void throttle_over(PairThrottle &t)
{
	int *n = t.n;
	if (!n[4]) n[8]++;
	else { n[8] = 0; n[5] = 0; n[12] = 0; }
	n[1]++;
	if (n[4] < 2 && n[29] > 0 && n[14] == 4) {
		if (n[31] < 2) { n[14] = 3; n[31]++; }
		else if (n[31] == 2) { n[14] = 0; n[15] = 0; n[31]++; }
	}
	if (n[14] == 5 && !n[35] && n[32] > 4 && n[32] < 8) { n[14] = 1; n[32]--; n[35]++; }
}

// a weak turn below the budget (image_processing.c:1517-1874): n1 creeps up, and a long script keyed on how many
// cycles have completed (n29, n32, n36, n28 ...) rewrites the mode n14 and the counters
NHW_HD void throttle_under(PairThrottle &t)
{
	int *n = t.n, *u = t.u;
	// (the walker is one thread on a latency chain: conditions are combined with & and | instead of && and ||, so the common
	// case runs through a handful of compares instead of a ladder of branches -- profiles/r02_notes.md)
	if (((n[1] == 6) & (u[8] == 0)) | (n[44] < -90000)) {
		if (n[1] == 6 && !u[8]) { n[1]++; u[8]++; n[44] = -100000; }
		else { n[1]++; u[8]++; n[44] = 0; }
	} else if (n[44] < 3) n[44]++;
	else { n[1] += 3; n[44] = 0; }
	if (!((n[29] > 0) & ((n[14] == 4) | (n[14] == 5) | (n[39] == 2) | (n[41] > 0)))) return;

	if (n[4] < 2 && n[1] == 15 && (n[14] == 4 || (n[14] == 5 && n[32] > 2))) {
		if (n[32] == 0 || n[32] == 2 || n[32] == 3 || (n[32] > 7 && n[32] < 500000)) {
			if (n[32] > 7 && n[14] == 5) { n[14] = 1; n[32] = 1000000; }
			else if (!n[34]) n[34] = 1;
			else { n[14] = 5; n[34] = 0; }
		}
		if (!n[32]) n[14] = 5;
		n[32]++;
	} else if (n[32] == 4 || n[32] == 5 || n[32] == 7) {
		if (n[37] == 4) n[14] = 3;
		else if (n[37] == 15) { n[14] = 3; n[32]++; }
		else if (n[32] == 7 && n[37] > -345000) {
			if (n[14] == 4) {
				if (!n[42]) n[37] -= 10000;
				if (n[38] > 0) {
					n[42]++;
					if (n[42] > 0 || (!n[42] && n[43] > 3)) {
						if (!n[42]) n[14] = n[43] == 14 ? 3 : n[43] == 24 ? 4 : 1;
						else n[14] = 1;
						n[39] = 0;
						if (n[42] > 5) { n[42] = -1; n[43]++; }
					} else if (n[42] == -1) { n[14] = 3; n[39] = 2; n[40] = -2; n[42] = 0; }
					else n[39] = 0;
				} else { n[14] = 5; n[39] = 1; n[42] = 0; }
			} else if (n[39] >= 1) {
				n[38]++;
				if (n[39] < 2) n[39] = (n[38] == 2 || n[38] == 4 || n[38] == 6 || n[38] == 9) ? 2 : 0;
				else {
					n[40]++;
					if (n[38] == 8) { n[39] = 0; n[40] = 0; }
					if (n[40] > 2) { n[40] = 0; n[39] = 0; }
				}
				if (n[38] >= 1 && n[38] <= 10) n[14] = 4;
			} else {
				n[40] = 1;
				if (n[38] == 1) n[39] = 2;
			}
		}
		if (n[37] >= 0) n[37]++;
	} else if (n[32] == 6 && n[36] < 118) {
		if (n[14] == 4 || n[14] == 5 || n[41] == 0 || n[41] > 3) n[36]++;
		if (n[41] > 3 && n[36] < 8) n[41] = 0;
		// script step -> (mode, what happens to n41: 0 cleared, 1 incremented, 4 set to 4)
		const int step[13] = {1, 2, 3, 4, 5, 6, 7, 8, 15, 31, 47, 100, 116};
		const int mode[13] = {1, 2, 1, 3, 3, 0, 2, 2, 1, 3, 2, 0, 2};
		const int op41[13] = {0, 0, 0, 0, 1, 0, 0, 4, 0, 1, 0, 1, 0};
		for (int k = 0; k < 13; k++)
			if (n[36] == step[k]) {
				n[14] = mode[k];
				if (op41[k] == 0) n[41] = 0;
				else if (op41[k] == 1) n[41]++;
				else n[41] = 4;
			}
	}

	if (n[28] < 14 && n[1] > 7) {
		if (n[14] == 5 && !n[28] && !n[33] && n[1] > 13 && n[31] > 0) { n[30] = 1; n[33] = 2; }
		else n[30]++;
		const int since = n[30] - n[33];    // weak turns since the script's clock was set
		if (!n[28] && since > 10 && n[33] > 0 && n[14] == 4) { n[14] = 3; n[15] += 6; n[28]++; }
		else if (n[28] == 1 && since > 70 && n[14] == 4 && n[1] == 11) { n[15] = 1; n[1] = 13; n[28]++; }
		else if (n[28] == 2 && n[31] > 2 && n[1] == 15 && n[15] > 1) { n[15] = 15; n[33] = n[30]; n[1] = 6; n[28]++; }
		else if (n[28] == 3 && since > 3 && n[31] > 2) { n[15] = 0; n[28]++; }
		else if (n[28] == 5 && since > 22 && n[31] > 2 && n[1] == 12) { n[15] = 3; n[1] = 9; n[28]++; }
		else if (n[28] == 4 && since > 6 && n[1] == 15) { n[14] = 1; n[15] += 6; n[1]++; n[28]++; }
		else if (n[28] == 6 && since > 54) { n[14] = 2; n[15] = 3; n[1] = 3; n[28]++; }
		else if (n[28] == 7 && since > 57) { n[14] = 2; n[15] = 8; n[1] = 8; n[28]++; }
		else if (n[28] == 8 && since > 84) { n[14] = 2; n[15] = 7; n[1] = 7; n[28]++; }
		else if (n[28] == 9 && since > 111) { n[14] = 2; n[15] = 3; n[1] = 7; n[28]++; }
		else if (n[28] == 10 && since > 116) { n[14] = 1; n[15] = 0; n[1] = 1; n[4] = 8; n[28]++; }
		else if (n[28] == 11 && since > 185) { n[14] = 0; n[15] = 4; n[1] = -17; n[28]++; }
		else if (n[28] == 12 && since > 187) { n[14] = 3; n[15] = 3; n[1] = -19; n[28]++; }
		else if (since == 9) { n[1] += (12 - n[4]) >> 2; n[4] = 10; }
		else if (n[28] > 0 && n[1] == 15 && u[1] < 11) {
			if (n[4] != 10) { if (u[1] == 4 || u[1] == 10) n[4] = 10; u[1]++; }
		} else if (n[28] == 13 && since > 188) { n[14] = 0; n[15] = 3; n[1] = -30; n[28]++; }
	}
}

// One pair of walk B at q <= 16 (image_processing.c:838-1924).  kA, kB: kernel values of the pair (may be rewritten:
// the rewritten values are what the later rules of the same pair see).  yA, yB: the two samples.  KA, KB: the
// stored kernel values (consumed ones are zeroed).  row = image row.
NHW_HD void throttle_pair(PairThrottle &t, const PreLowParams &p, int row, int &kA, int &kB, int16_t &yA, int16_t &yB,
                          int16_t &KA, int16_t &KB)
{
	int *n = t.n;
	const int sh = p.sharp, s2 = p.sharp2;
	const bool hitA = nhw_iabs(kA) > sh, hitB = nhw_iabs(kB) > sh;
	if (!n[1]) {
		// ---- strong turn
		n[2] = 0;
		if (hitA) {
			yA = (int16_t)(yA + (kA > 0 ? 2 : -2));
			if (nhw_iabs(kB) > s2 || n[8] == 1) {
				KA = 0;
				if ((n[19] < 262144 || (n[20] >= 3 && n[20] < 262144)) && nhw_iabs(kA) > sh + 96 && n[6] > 0 && row > 2) {
					if (n[20] >= 3 && n[19] >= 524288) { n[6] = 7000000; n[20] = 524288; }
					if (n[19] > 0 && n[19] < 262144) {
						if (n[20] > 2 || (n[20] == 2 && n[6] > 3 && !n[23]) || (n[20] == 2 && n[6] > 14 && n[23] > 0)) {
							if (n[23] == 1) n[6] = 5000000;
							n[23]++; n[21]++;
							if (n[21] >= 2) n[19] = 524288;
						}
					}
					if (!n[19]) { n[6]++; n[20] = 1; }
					n[19]++;
				}
			}
			n[2] = 1;
		}
		if (hitB) {
			const int step2 = kB > 0 ? 2 : -2;
			if ((n[2] == 1 || n[12] == 1) && (!n[14] || n[14] == 4 || n[14] == 5) && !n[3] && n[2] == 1) {
				if (nhw_iabs(kA) > 3000) kA = kA > 0 ? s2 + 5 : -s2 - 5;       // markers count as just-over-threshold values
				if (nhw_iabs(kB) > 3000) kB = kB > 0 ? s2 + 22 : -s2 - 22;
				if (nhw_iabs(kA) < (nhw_iabs(kB) >> 2)) {
					yA = (int16_t)(yA + (kA > 0 ? -1 : 1));
					KA = (int16_t)kA;
					yB = (int16_t)(yB + (kB > 0 ? 2 : -2));
					if (nhw_iabs(kA) > s2) KB = 0;
				} else yB = (int16_t)(yB + (kB > 0 ? 1 : -1));
				n[3] = 1;
			} else {
				const bool cyc = (n[2] == 1 || n[12] == 1) && (!n[14] || n[14] == 4 || n[14] == 5);
				yB = (int16_t)(yB + step2);
				if (nhw_iabs(kA) > s2) KB = 0;
				if (cyc) n[3] = n[3] == 1 ? 2 : n[3] == 2 ? 3 : 0;
			}
			if (n[14] == 2) { n[14] = 1; n[26] = 3; if (n[25] > 0) n[25]++; }
			if (n[14] == 1) {
				if (n[26] < 4) n[26]++;
				else { n[14] = 2; n[26] = 0; }
			}
		}
		if (nhw_iabs(kA) > sh || nhw_iabs(kB) > sh) n[13] = 1;
		if (n[14] == 1 || n[14] == 2) n[27]++;
		else n[27] = 0;
		if (n[27] > 2) n[14] = 1;
		if (n[14] == 1) {
			n[14] = 4;
			if (!n[25]) { n[15]++; n[25] = 1; }
			else { n[25]++; if (n[25] > 3) n[25] = 0; }
		}
		n[1] = 1;
		return;
	}
	// ---- weak turn
	if (hitA | hitB) {
		if (hitA) { yA = (int16_t)(yA + (kA > 0 ? 1 : -1)); n[1]++; n[4]++; }
		if (hitB) { yB = (int16_t)(yB + (kB > 0 ? 1 : -1)); n[1]++; n[4]++; }
	}
	bool spent;                                  // n17: the weak-turn budget of this cycle is used up
	if (n[4] < 10) spent = (n[4] == n[10]) & (n[1] == n[11]);
	else if ((n[4] > 10) | (n[1] != 15)) {
		if (!n[18]) { spent = true; n[18] = 1; }
		else { spent = false; n[18]++; if (n[18] > 15) n[18] = 0; }
	} else spent = (n[4] == n[10]) & (n[1] == n[11]);
	n[17] = spent ? 1 : 0;
	if (n[6] > 4000000) {
		if (n[6] > 6000000) { n[6] = 0; n[22] = 0; }
		else { n[6] = 0; n[22] = n[21] == 1 ? 1 : 0; }
	}
	if (spent | (n[1] > 2000003)) throttle_rearm(t);
	else if (n[1] >= 15) throttle_over(t);
	else throttle_under(t);
	if ((n[8] > 6) & (n[4] == 0) & (n[1] > 1) & (n[1] < 15)) {
		n[5]++;
		if (n[5] < 35) {
			n[1] = 0;
			if (!n[13]) { n[12] = 1; n[13] = 1; }
			else { n[12] = 0; n[13]++; if (n[13] > 3) n[13] = 0; }
		} else n[12] = 0;
	}
	if ((n[1] > 15) & (n[1] < 1000000)) { n[1] = 0; n[4] = 0; n[29]++; }
}

// ---- the quiet stretch.  At most qualities nine pairs in ten have no value above the threshold, and the throttle then
// runs on its own: a weak turn below the budget in which only n1, n44 (and once u8) move, until n1 passes 15 and the cycle
// restarts.  Everything else throttle_pair looks at on that path (n4, n6, n8, n10, n11, n14, n29, n39, n41) stands still,
// so it is summed up once in `steady` -- recomputed whenever the general path or a cycle restart has run -- and a quiet
// pair then costs a dozen instructions instead of ~150.  throttle_quiet_step is exactly throttle_pair restricted to that
// path (no hit, weak turn, budget not spent, n1 < 15, throttle_under's early return); it returns false, having changed
// nothing, when the pair is not on it.
NHW_HD bool throttle_steady(const PairThrottle &t)
{
	const int *n = t.n;
	const bool busy = (n[29] > 0) & ((n[14] == 4) | (n[14] == 5) | (n[39] == 2) | (n[41] > 0));
	return (n[4] < 10) & (n[6] <= 4000000) & !busy & !((n[8] > 6) & (n[4] == 0));
}
NHW_HD bool throttle_quiet_step(PairThrottle &t, bool &steady)
{
	int *n = t.n, *u = t.u;
	const int n1 = n[1];
	if (!((n1 != 0) & (n1 < 15) & !((n[4] == n[10]) & (n1 == n[11])))) return false;
	n[17] = 0;
	if (((n1 == 6) & (u[8] == 0)) | (n[44] < -90000)) {
		if (n1 == 6 && !u[8]) { n[1]++; u[8]++; n[44] = -100000; }
		else { n[1]++; u[8]++; n[44] = 0; }
	} else if (n[44] < 3) n[44]++;
	else { n[1] += 3; n[44] = 0; }
	if (n[1] > 15) { n[1] = 0; n[4] = 0; n[29]++; steady = throttle_steady(t); }   // (n1 < 1000000 here)
	return true;
}

// the q <= 14 smoothing of walk B for one pixel: O = the plane before the stage, kv = the pixel's kernel value after
// walk A.  Pointwise (it reads O only), so it is applied to a whole row before the row's pairs are walked.
NHW_HD bool pre_low_smooth_cell(const int16_t *O, int at, int kv, const PreLowParams &p, int &out)
{
	if (!(nhw_iabs(kv) > 4 && nhw_iabs(kv) < p.smooth_below)) return false;
	const int up = O[at - PW], lf = O[at - 1], dn = O[at + PW], rt = O[at + 1];
	if (!(nhw_iabs(up - lf) < 4 && nhw_iabs(lf - dn) < 4 && nhw_iabs(dn - rt) < 4 && nhw_iabs(rt - up) < 4)) return false;
	out = ((O[at] << 2) + lf + rt + up + dn + 4) >> 3;
	return true;
}

// One row of walk B.  Yr, Kr, Mr point at column 0 of the row (the plane itself or a staged copy); the row's smoothing
// has been applied to Yr already.  `a` = the mid-range rule's one-pair memory, carried from row to row like the throttle.
NHW_HD void pre_low_walk_b_row(PairThrottle &t, int &a, const PreLowParams &p, int r, int16_t *Yr, int16_t *Kr, uint8_t *Mr)
{
	const int sh = p.sharp, small_below = sh < 22 ? sh : 22;
	bool steady = throttle_steady(t);
	for (int j = 1; j < 510; j += 2) {
		// the pair is columns (j, j + 1)
		int kA = Kr[j], kB = Kr[j + 1];
		// a pair whose two values are both at most min(sharp, 22) in magnitude is below every rule of this walk: no hit, not
		// an opposite-sign pair (both need > sharp), not a mid-range pair (one value would have to reach 23) -- only the
		// throttle moves, and the mid-range rule forgets its one-pair memory
		const int big = nhw_iabs(kA) > nhw_iabs(kB) ? nhw_iabs(kA) : nhw_iabs(kB);
		{
			// the commonest pair of all, as ONE branch (a lone thread pays tens of cycles for every branch it takes): small
			// pair, steady throttle, weak turn below the budget, the plain n44 / n1 tick of throttle_under
			int *n = t.n;
			const int n1 = n[1], n44 = n[44];
			const bool tick = (big <= small_below) & steady & (n1 != 0) & (n1 < 15) & !((n[4] == n[10]) & (n1 == n[11])) &
			                  !(((n1 == 6) & (t.u[8] == 0)) | (n44 < -90000));
			if (tick) {
				const bool low = n44 < 3;
				const int m1 = n1 + (low ? 0 : 3);
				n[44] = low ? n44 + 1 : 0;
				n[17] = 0;
				a = 0;
				if (m1 > 15) { n[1] = 0; n[4] = 0; n[29]++; steady = throttle_steady(t); }
				else n[1] = m1;
				continue;
			}
		}
		if ((big <= small_below) & steady) {
			if (throttle_quiet_step(t, steady)) { a = 0; continue; }
		}
		const bool quiet = big <= sh;
		if (!(quiet && steady && throttle_quiet_step(t, steady))) {
			throttle_pair(t, p, r, kA, kB, Yr[j], Yr[j + 1], Kr[j], Kr[j + 1]);
			steady = throttle_steady(t);
		}
		// opposite-sign pair just above the threshold: push them apart, and remember which way (M)
		if ((nhw_iabs(kA) > sh) & (nhw_iabs(kA) <= sh + 20) & (nhw_iabs(kB) > sh) & (nhw_iabs(kB) <= sh + 20)) {
			if (kA > 0 && kB < 0) { Yr[j]++; Yr[j + 1]--; Mr[j] = 2; Mr[j + 1] = 3; }
			else if (kA < 0 && kB > 0) { Yr[j]--; Yr[j + 1]++; Mr[j] = 3; Mr[j + 1] = 2; }
		}
		if (p.midrange_on) {
			// the 10..32 / >= 23 rule of the q > 16 path, without its 176 / 201 part
			int d0 = 0, d1 = 0;
			const int ar = nhw_iabs(kA), ac = nhw_iabs(kB);
			if ((ar > 10) & (ar < 32) & (ac >= 23)) {
				const int sg = kA > 0 ? 1 : -1;
				if (ar < 16) { if (kB * sg > 0 && ac < 32 && ar > 11) d1 = sg; d0 = sg; }
				else d0 = a ? sg : 2 * sg;
				a = 0;
			} else {
				a = 0;
				if ((ac > 10) & (ac < 32) & (ar >= 23)) {
					const int sg = kB > 0 ? 1 : -1;
					if (ac < 16) { if (kA * sg > 0 && ar < 32 && ac > 11) d0 = sg; d1 = sg; }
					else { d1 = 2 * sg; a = 1; }
				}
			}
			if (d0 | d1) {
				Yr[j] = (int16_t)(Yr[j] + d0);
				Yr[j + 1] = (int16_t)(Yr[j + 1] + d1);
			}
		}
	}
}

NHW_HDN void pre_low_walk_b(int16_t *Y, const int16_t *O, int16_t *K, uint8_t *M, const PreLowParams &p)
{
	PairThrottle t;
	t.init();
	int a = 0;
	for (int r = 1; r < 511; r++) {
		if (p.smooth_on)
			for (int j = 1; j < 511; j++) {
				int v;
				if (pre_low_smooth_cell(O, r * PW + j, K[r * PW + j], p, v)) Y[r * PW + j] = (int16_t)v;
			}
		pre_low_walk_b_row(t, a, p, r, Y + r * PW, K + r * PW, M + r * PW);
	}
}

// ---- walk C -----------------------------------------------------------------------------------------------------
struct PreWalkC { int skip_first, gate, pa, na, pb, nb; };   // image-wide toggles (t1..t6 of this loop)

// a marker becomes 0 on its first appearance and +-5000 on the next two (cycle of three); 7000 becomes sharp2 + 22
NHW_HD void walk_c_resolve(int16_t &cell, int v, int &pos_cycle, int &neg_cycle, int s2)
{
	if (v == 20000) {
		if (!pos_cycle) { cell = 0; pos_cycle = 1; }
		else { cell = 5000; pos_cycle = pos_cycle == 1 ? 2 : 0; }
	} else if (v == -20000) {
		if (!neg_cycle) { cell = 0; neg_cycle = 1; }
		else { cell = -5000; neg_cycle = neg_cycle == 1 ? 2 : 0; }
	} else if (v == 7000) cell = (int16_t)(s2 + 22);
}

// One row of walk C.  Y, K, M point at column 0 of row r - 1 of a window that holds rows r - 1 and r back to back
// (the plane itself, or a staged copy of the two rows): every cell the row can touch lies in that window.
NHW_HD void pre_low_walk_c_row(PreWalkC &w, const PreLowParams &p, int r, int16_t *Y, int16_t *K, uint8_t *M)
{
	const int sh = p.sharp, s2 = p.sharp2;
	auto touch = [&](int at, int d) { Y[at] = (int16_t)(Y[at] + d); M[at] = 1; };
	int misses = 0, back = 0, fresh = 0;     // per-row cursor state (e, t, f)
	for (int j = 1, s = PW + 1; j < 509; j++, s++) {
		int kA = K[s];
		j++; s++;
		int kB = K[s];
		if ((nhw_iabs(kA) > 6000) | (nhw_iabs(kB) > 6000)) {
			if (nhw_iabs(kA) > 6000) {
				walk_c_resolve(K[s - 1], kA, w.pa, w.na, s2);
				if (!w.gate) { walk_c_resolve(K[s], kB, w.pb, w.nb, s2); w.gate = 1; }
				else w.gate = 0;
				if (!w.skip_first) { w.skip_first = 1; continue; }
				w.skip_first = 0;
			} else {
				walk_c_resolve(K[s], kB, w.pb, w.nb, s2);
				continue;
			}
		}
		// strong value next to a weak one (the stale kA / kB are used on purpose: a resolved marker still counts
		// with its marker value here)
		const bool strongA = (nhw_iabs(kA) > sh + 20) & (nhw_iabs(kB) > (sh >> 1)) & (nhw_iabs(kB) <= s2);
		const bool strongB = !strongA & (nhw_iabs(kB) > sh + 20) & (nhw_iabs(kA) > (sh >> 1)) & (nhw_iabs(kA) <= s2);
		if (strongA | strongB) {
			const int big = strongA ? kA : kB, small = strongA ? kB : kA;
			const int at_big = strongA ? s - 1 : s, at_small = strongA ? s : s - 1;
			if (big != 0) {
				const int sg = big > 0 ? 1 : -1;
				touch(at_big, sg);
				if (small * sg > 0) touch(at_small, 2 * sg);
				if (r * PW + j >= 2 * PW + 2) {
					// the two cells of the row above that sit over the pair
					const int hi = s - PW, lo = s - PW - 1;
					const int k_hi = K[hi], k_lo = K[lo];
					if (strongA) {
						if (k_hi * sg > 4) touch(hi, sg);
						if (k_lo * sg > 4) touch(lo, sg);
						if (k_hi * sg < -24 && !back) touch(hi, -sg);
						if (k_lo * sg < -24 && !back) touch(lo, -sg);
					} else {
						if (k_lo * sg > 4) touch(lo, sg);
						if (k_hi * sg > 4) touch(hi, sg);
						if (k_lo * sg < -24 && !back) touch(lo, -sg);
						if (k_hi * sg < -24 && !back) touch(hi, -sg);
					}
				}
				misses = 0; fresh = 0;
			}
			if (back == 1) { j++; s++; back = 0; }
			else if (back == 2) { j += 3; s += 3; back = 0; }
		} else {
			misses++;
			if (!back) fresh++;
			if (misses == 2) { j -= 3; s -= 3; misses = 0; back = 1; }
			else if (back == 1) {
				j++; s++; back = 0; misses = 0;
				if (fresh == 4) {
					if (nhw_iabs(K[s - 5]) <= s2 || nhw_iabs(K[s - 2]) <= s2) { j -= 5; s -= 5; back = 2; }
					fresh = 0;
				}
			} else if (back == 2) { j += 3; s += 3; back = 0; misses = 0; fresh = 0; }
		}
	}
}

NHW_HDN void pre_low_walk_c(int16_t *Y, int16_t *K, uint8_t *M, const PreLowParams &p)
{
	PreWalkC w = {0, 0, 0, 0, 0, 0};
	for (int r = 1; r < 511; r++) pre_low_walk_c_row(w, p, r, Y + (r - 1) * PW, K + (r - 1) * PW, M + (r - 1) * PW);
}

// ---- walk D -----------------------------------------------------------------------------------------------------
NHW_HD void pre_low_walk_d_row(int16_t *Y, const int16_t *K, const uint8_t *M, const PreLowParams &p, int r)
{
	const int sh = p.sharp, s2 = p.sharp2;
	auto near_thr = [&](int v) { return nhw_iabs(v) > sh && nhw_iabs(v) <= sh + 20; };
	auto near_thr2 = [&](int v) { return nhw_iabs(v) > s2 && nhw_iabs(v) <= s2 + 20; };
	for (int j = 1, s = r * PW + 1; j < 510; j++, s++) {
		const int kA = K[s];
		j++; s++;
		const int kB = K[s];
		if (nhw_iabs(kA) > 4000 || nhw_iabs(kB) > 4000) continue;
		const int mA = M[s - 1], mB = M[s];
		bool back = false;     // re-pair starting from the pair's second pixel
		if (near_thr(kA) && near_thr(kB)) {
			bool moved = false;
			if (mA != 1 && mB != 1) {
				if (kA > 0 && kB > 0) {
					moved = true;
					const bool first = kA >= kB;          // the larger one is raised, unless walk B raised it already
					if (first ? mA != 2 : mB != 2) Y[first ? s - 1 : s]++;
					else if (first ? mB != 2 : mA != 2) Y[first ? s : s - 1]++;
				} else if (kA < 0 && kB < 0) {
					moved = true;
					const bool first = kA <= kB;
					if (first ? mA != 3 : mB != 3) Y[first ? s - 1 : s]--;
					else if (first ? mB != 3 : mA != 3) Y[first ? s : s - 1]--;
				}
			}
			if (!moved && j < 508 && near_thr(K[s + 1])) back = (kB > 0 && K[s + 1] > 0) || (kB < 0 && K[s + 1] < 0);
		} else if (nhw_iabs(kA) > sh + 56 && nhw_iabs(kB) > sh + 56) {
			if (!mA && !mB) {
				if (kA > 0 && kB < 0) { Y[s - 1]++; Y[s]--; }
				else if (kA < 0 && kB > 0) { Y[s - 1]--; Y[s]++; }
				else if (nhw_iabs(kA) > sh + 96 && nhw_iabs(kB) > sh + 96) {
					if (kA > 0 && kB > 0) Y[kA > kB ? s - 1 : s]++;
					else if (kA < 0 && kB < 0) Y[kA < kB ? s - 1 : s]--;
				}
			}
		} else if (nhw_iabs(kA) > sh + 160 && near_thr2(kB)) {
			if (!mA && !mB) {
				if (kA > 0 && kB > 0) Y[s]--;
				else if (kA < 0 && kB < 0) Y[s]++;
				else back = j < 506 && nhw_iabs(K[s + 1]) > sh + 160 && nhw_iabs(K[s + 2]) <= s2;
			} else back = j < 506 && nhw_iabs(K[s + 1]) > sh + 160 && nhw_iabs(K[s + 2]) > s2 + 20;
		} else if (nhw_iabs(kB) > sh + 160 && near_thr2(kA)) {
			if (!mA && !mB) {
				if (kA > 0 && kB > 0) Y[s - 1]--;
				else if (kA < 0 && kB < 0) Y[s - 1]++;
				else back = j < 508 && near_thr2(K[s + 1]);
			} else back = true;
		} else back = true;
		if (back) { j--; s--; }
	}
}
