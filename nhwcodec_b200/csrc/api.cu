// api.cu -- context management and the C-ABI entry points of libnhw_cuda.so
// (declared in include/nhw_cuda.h).  There is no CPU fallback anywhere in this library:
// without a usable CUDA device nhw_create() fails and every other call needs a context.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/nhw_cuda.h"
#include "nhw_ctx.h"
#include "nhw_dev.cuh"
#include "dec_parse.h"

namespace nhw {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
}

bool check(cudaError_t e, const char *what)
{
	if (e == cudaSuccess) return true;
	set_error("%s: %s", what, cudaGetErrorString(e));
	return false;
}

}  // namespace nhw

using nhw::check;

namespace nhw {

struct ProfEntry { const char *label; cudaEvent_t a, b; };
struct ProfState {
	std::vector<ProfEntry> open;                 // launched, not yet resolved
	std::vector<cudaEvent_t> pool;
	std::map<std::string, std::pair<double, uint64_t>> acc;   // label -> (ms, launches)
};

static cudaEvent_t prof_event(ProfState *p)
{
	if (!p->pool.empty()) { cudaEvent_t e = p->pool.back(); p->pool.pop_back(); return e; }
	cudaEvent_t e;
	cudaEventCreate(&e);
	return e;
}

// debug bisection aid: true = do not launch (a stop point was reached earlier in this call)
bool dbg_skip(nhw_ctx *c, const char *label)
{
	if (!c->dbg_label[0]) return false;
	if (c->dbg_stopped) return true;
	if (strcmp(label, c->dbg_label) == 0 && ++c->dbg_seen >= c->dbg_count) c->dbg_stopped = 2;   // launch this one, then stop
	if (c->dbg_stopped == 2) { c->dbg_stopped = 1; return false; }
	return false;
}

void prof_begin(nhw_ctx *c, const char *label)
{
	if (!c->profile) return;
	ProfState *p = static_cast<ProfState *>(c->prof);
	ProfEntry en{label, prof_event(p), prof_event(p)};
	cudaEventRecord(en.a, c->stream);
	p->open.push_back(en);
}

void prof_end(nhw_ctx *c)
{
	if (!c->profile) return;
	ProfState *p = static_cast<ProfState *>(c->prof);
	cudaEventRecord(p->open.back().b, c->stream);
}

static void prof_resolve(nhw_ctx *c)
{
	ProfState *p = static_cast<ProfState *>(c->prof);
	if (!p) return;
	for (auto &en : p->open) {
		float ms = 0.f;
		if (cudaEventSynchronize(en.b) == cudaSuccess && cudaEventElapsedTime(&ms, en.a, en.b) == cudaSuccess) {
			auto &slot = p->acc[en.label];
			slot.first += ms;
			slot.second += 1;
		}
		p->pool.push_back(en.a);
		p->pool.push_back(en.b);
	}
	p->open.clear();
}

}  // namespace nhw

static int env_int(const char *name, int dflt, int lo, int hi)
{
	const char *e = getenv(name);
	const int v = e ? atoi(e) : dflt;
	return v < lo ? lo : v > hi ? hi : v;
}

template <typename T>
static bool dev_alloc0(T **p, size_t count)
{
	if (!check(cudaMalloc((void **)p, count * sizeof(T)), "cudaMalloc")) return false;
	return check(cudaMemset(*p, 0, count * sizeof(T)), "cudaMemset");
}

extern "C" {

const char *nhw_last_error(void) { return nhw::g_err; }
int nhw_version(void) { return 100; }

int nhw_create(int device, int max_batch, nhw_ctx **out)
{
	if (!out || max_batch <= 0) return NHW_ERR_ARG;
	*out = nullptr;
	int count = 0;
	if (!check(cudaGetDeviceCount(&count), "cudaGetDeviceCount") || count <= 0) {
		if (count <= 0 && nhw::g_err[0] == 0) nhw::set_error("no CUDA device");
		return NHW_ERR_CUDA;
	}
	if (device < 0 || device >= count) { nhw::set_error("device %d out of range (%d devices)", device, count); return NHW_ERR_ARG; }
	if (!check(cudaSetDevice(device), "cudaSetDevice")) return NHW_ERR_CUDA;
	nhw_ctx *c = new (std::nothrow) nhw_ctx();
	if (!c) return NHW_ERR_NOMEM;
	memset(c, 0, sizeof *c);
	c->device = device;
	c->max_batch = max_batch;
	c->prof = new nhw::ProfState();
	const size_t B = (size_t)max_batch;
	bool ok = true;
	for (int k = 0; k < NHW_LANES && ok; k++) {
		ok = ok && check(cudaStreamCreateWithFlags(&c->lanes[k], cudaStreamNonBlocking), "cudaStreamCreate");
		ok = ok && check(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming), "cudaEventCreate");
	}
	ok = ok && check(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming), "cudaEventCreate");
	ok = ok && check(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
	ok = ok && check(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking), "cudaStreamCreate");
	ok = ok && check(cudaStreamCreateWithFlags(&c->chroma_stream, cudaStreamNonBlocking), "cudaStreamCreate");
	ok = ok && check(cudaEventCreateWithFlags(&c->ev_chroma0, cudaEventDisableTiming), "cudaEventCreate");
	ok = ok && check(cudaEventCreateWithFlags(&c->ev_chroma1, cudaEventDisableTiming), "cudaEventCreate");
	for (int k = 0; k < NHW_MAX_SUB && ok; k++) {
		ok = ok && check(cudaEventCreateWithFlags(&c->ev_sub[k], cudaEventDisableTiming), "cudaEventCreate");
		ok = ok && check(cudaEventCreateWithFlags(&c->ev_up[k], cudaEventDisableTiming), "cudaEventCreate");
	}

	c->stream = c->lanes[0];
	{
		int sms = 148;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
		NhwTuning &t = c->tune;
		t.lanes_device = env_int("NHW_LANES_DEVICE", 1, 1, NHW_LANES);
		t.chroma_stream = env_int("NHW_CHROMA_STREAM", 1, 0, 1);
		t.subs_encode = env_int("NHW_SUBS_ENCODE", 16, 1, NHW_MAX_SUB);
		t.lanes_encode = env_int("NHW_LANES_ENCODE", 4, 1, NHW_LANES);
		t.subs_decode = env_int("NHW_SUBS_DECODE", 8, 1, NHW_MAX_SUB);
		t.lanes_decode = env_int("NHW_LANES_DECODE", 4, 1, NHW_LANES);
		t.dsf_streams = env_int("NHW_DSF_STREAMS", 0, 0, 32);   // 0: chosen per launch from the number of streams
		while (t.dsf_streams && 32 % t.dsf_streams) t.dsf_streams--;
		t.dsf_job_mask = env_int("NHW_DSF_JOBS", 15, 0, 15);
		t.rows_grid_cap = sms * env_int("NHW_ROWS_CTAS_PER_SM", 24, 1, 64);
		t.fetch_kernel = env_int("NHW_FETCH_KERNEL", 1, 0, 1);
		t.plw_phases = env_int("NHW_PLW_PHASES", 15, 0, 15);
	}
	ok = ok && nhw::front_device_init(c) && nhw::encode_device_init(c) && nhw::decode_device_init(c);
	// every workspace array is zero-filled once: guard bands and never-written borders must
	// read as 0 (canonical oracle semantics)
	ok = ok && dev_alloc0(&c->rgb, B * NHW_RGB_BYTES);
	ok = ok && dev_alloc0(&c->y_jpeg, B * NHW_Y_SLOT) && dev_alloc0(&c->y_proc, B * NHW_Y_SLOT);
	ok = ok && dev_alloc0(&c->y_aux, B * NHW_Y_SLOT) && dev_alloc0(&c->y_aux2, B * NHW_Y_SLOT);
	ok = ok && dev_alloc0(&c->y_ll1, B * NHW_C_SLOT) && dev_alloc0(&c->y_ll2save, B * NHW_C_SLOT);
	ok = ok && dev_alloc0(&c->c_u8, B * 2 * NHW_CPLANE);
	ok = ok && dev_alloc0(&c->c_jpeg, B * 2 * NHW_C_SLOT) && dev_alloc0(&c->c_proc, B * 2 * NHW_C_SLOT);
	ok = ok && dev_alloc0(&c->c_aux, B * 2 * NHW_C_SLOT);
	ok = ok && dev_alloc0(&c->c_ll1, B * 2 * NHW_Q_SLOT) && dev_alloc0(&c->c_ll2save, B * 2 * NHW_Q_SLOT);
	ok = ok && dev_alloc0(&c->rowmap, B * 512) && dev_alloc0(&c->rowcarry, B * 512);
	ok = ok && dev_alloc0(&c->enc_bytes, B * (size_t)ENC_BYTES_SLOT) && dev_alloc0(&c->enc_hdr, B);
	ok = ok && dev_alloc0(&c->dec_bytes, B * (size_t)NHW_DEC_BYTES_SLOT);
	ok = ok && dev_alloc0(&c->out_dev, B * (size_t)NHW_MAX_STREAM_BYTES) && dev_alloc0(&c->pack_dev, B * (size_t)NHW_MAX_STREAM_BYTES);
	ok = ok && dev_alloc0(&c->len_dev, B) && dev_alloc0(&c->status_dev, B) && dev_alloc0(&c->offs_dev, B + 1 + NHW_MAX_SUB);
	ok = ok && dev_alloc0(&c->dec_yuv, B * (size_t)NHW_RGB_BYTES);
	ok = ok && check(cudaMalloc(&c->dec_desc_dev, B * sizeof(DecDesc)), "cudaMalloc");
	ok = ok && check(cudaMallocHost(&c->dec_desc_host, B * sizeof(DecDesc)), "cudaMallocHost");
	ok = ok && check(cudaMallocHost((void **)&c->offs_host, (B + 1 + NHW_MAX_SUB) * sizeof(uint64_t)), "cudaMallocHost");
	ok = ok && check(cudaMallocHost((void **)&c->status_host, B * sizeof(int32_t)), "cudaMallocHost");
	if (!ok || !check(cudaDeviceSynchronize(), "nhw_create sync")) { nhw_destroy(c); return NHW_ERR_CUDA; }
	*out = c;
	return NHW_OK;
}

void nhw_destroy(nhw_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	void *ptrs[] = {c->rgb, c->y_jpeg, c->y_proc, c->y_aux, c->y_aux2, c->y_ll1, c->y_ll2save, c->c_u8, c->c_jpeg,
	                c->c_proc, c->c_aux, c->c_ll1, c->c_ll2save, c->rowmap, c->rowcarry, c->enc_bytes, c->dec_bytes, c->enc_hdr,
	                c->out_dev, c->pack_dev, c->len_dev, c->status_dev, c->offs_dev};
	for (void *p : ptrs)
		if (p) cudaFree(p);
	if (c->dec_yuv) cudaFree(c->dec_yuv);
	if (c->dec_desc_dev) cudaFree(c->dec_desc_dev);
	if (c->dec_desc_host) cudaFreeHost(c->dec_desc_host);
	if (c->offs_host) cudaFreeHost(c->offs_host);
	if (c->status_host) cudaFreeHost(c->status_host);
	for (int k = 0; k < NHW_LANES; k++) {
		if (c->lanes[k]) cudaStreamDestroy(c->lanes[k]);
		if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
	}
	if (c->ev_fork) cudaEventDestroy(c->ev_fork);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->up_stream) cudaStreamDestroy(c->up_stream);
	if (c->chroma_stream) cudaStreamDestroy(c->chroma_stream);
	if (c->ev_chroma0) cudaEventDestroy(c->ev_chroma0);
	if (c->ev_chroma1) cudaEventDestroy(c->ev_chroma1);
	for (int k = 0; k < NHW_MAX_SUB; k++) {
		if (c->ev_sub[k]) cudaEventDestroy(c->ev_sub[k]);
		if (c->ev_up[k]) cudaEventDestroy(c->ev_up[k]);
	}

	if (c->prof) {
		nhw::ProfState *p = static_cast<nhw::ProfState *>(c->prof);
		nhw::prof_resolve(c);
		for (cudaEvent_t e : p->pool) cudaEventDestroy(e);
		delete p;
	}
	delete c;
}

uint64_t nhw_launch_count(const nhw_ctx *c) { return c ? c->launches : 0; }

void *nhw_stream(const nhw_ctx *c) { return c ? (void *)c->stream : nullptr; }

int nhw_debug_stop_after(nhw_ctx *c, const char *label, int occurrence)
{
	if (!c) return NHW_ERR_ARG;
	c->dbg_label[0] = 0;
	c->dbg_seen = c->dbg_stopped = 0;
	c->dbg_count = occurrence > 0 ? occurrence : 1;
	if (label) { strncpy(c->dbg_label, label, sizeof c->dbg_label - 1); c->dbg_label[sizeof c->dbg_label - 1] = 0; }
	return NHW_OK;
}

int nhw_debug_read(nhw_ctx *c, const char *what, int img, void *host, size_t bytes)
{
	if (!c || !what || !host || img < 0 || img >= c->max_batch) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	const void *src = nullptr;
	const size_t i = (size_t)img;
	if (!strcmp(what, "proc")) src = c->y_proc + NHW_GUARD_S + i * NHW_Y_SLOT;
	else if (!strcmp(what, "jpeg")) src = c->y_jpeg + NHW_GUARD_S + i * NHW_Y_SLOT;
	else if (!strcmp(what, "aux")) src = c->y_aux + NHW_GUARD_S + i * NHW_Y_SLOT;
	else if (!strcmp(what, "ll1")) src = c->y_ll1 + NHW_GUARD_S + i * NHW_C_SLOT;
	else if (!strcmp(what, "ll2s")) src = c->y_ll2save + NHW_GUARD_S + i * NHW_C_SLOT;
	else if (!strcmp(what, "cproc_u")) src = c->c_proc + NHW_GUARD_S + (2 * i) * NHW_C_SLOT;
	else if (!strcmp(what, "cproc_v")) src = c->c_proc + NHW_GUARD_S + (2 * i + 1) * NHW_C_SLOT;
	else if (!strcmp(what, "cjpeg_u")) src = c->c_jpeg + NHW_GUARD_S + (2 * i) * NHW_C_SLOT;
	else if (!strcmp(what, "cjpeg_v")) src = c->c_jpeg + NHW_GUARD_S + (2 * i + 1) * NHW_C_SLOT;
	else if (!strcmp(what, "scan")) src = c->enc_bytes + i * (size_t)ENC_BYTES_SLOT + OFF_SCAN;
	else if (!strcmp(what, "tree1")) src = c->enc_bytes + i * (size_t)ENC_BYTES_SLOT + OFF_TREE1;
	else if (!strcmp(what, "llcode")) src = c->enc_bytes + i * (size_t)ENC_BYTES_SLOT + OFF_LLCODE;
	else if (!strcmp(what, "hdr")) src = c->enc_hdr + i;
	else return NHW_ERR_ARG;
	cudaStreamSynchronize(c->stream);
	return check(cudaMemcpy(host, src, bytes, cudaMemcpyDeviceToHost), "nhw_debug_read") ? NHW_OK : NHW_ERR_CUDA;
}

long nhw_debug_color_check(nhw_ctx *c)
{
	if (!c) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	return nhw::color_fast_path_mismatches(c);
}

long nhw_debug_dec_color_check(nhw_ctx *c)
{
	if (!c) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	return nhw::dec_color_fast_path_mismatches(c);
}

int nhw_profile(nhw_ctx *c, int enable)
{
	if (!c) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	nhw::prof_resolve(c);
	if (enable == 2) static_cast<nhw::ProfState *>(c->prof)->acc.clear();   // reset counters, keep on
	c->profile = enable ? 1 : 0;
	return NHW_OK;
}

long nhw_profile_read(nhw_ctx *c, char *buf, size_t cap)
{
	if (!c || !buf || cap == 0) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	nhw::prof_resolve(c);
	nhw::ProfState *p = static_cast<nhw::ProfState *>(c->prof);
	size_t off = 0;
	for (auto &kv : p->acc) {
		int w = snprintf(buf + off, cap - off, "%s\t%.6f\t%llu\n", kv.first.c_str(), kv.second.first,
		                 (unsigned long long)kv.second.second);
		if (w < 0 || (size_t)w >= cap - off) break;
		off += (size_t)w;
	}
	buf[off < cap ? off : cap - 1] = 0;
	return (long)off;
}

static int finish(nhw_ctx *c, const char *what)
{
	if (!check(cudaGetLastError(), what)) return NHW_ERR_CUDA;
	if (!check(cudaStreamSynchronize(c->stream), what)) return NHW_ERR_CUDA;
	return NHW_OK;
}

// ---- lanes: sub-chunks side by side on their own streams and workspace slices --------------------
struct LanePlan {
	int count;           // images this wave takes (<= subs * slot <= max_batch)
	int subs;            // sub-chunks in this wave (each on its own workspace slice)
	int streams;         // streams they are dealt to, round robin
	int slot;            // workspace images per sub-chunk
	int first[NHW_MAX_SUB + 1];   // image range of each sub-chunk inside the wave
};

// Split the next `m` images (m <= max_batch) into sub-chunks.  Per-kernel profiling and the debug stop are
// defined on one stream, so they run as one sub-chunk.
static LanePlan plan_lanes(const nhw_ctx *c, int m, int want_subs, int want_streams)
{
	LanePlan p;
	int subs = (c->profile || c->dbg_label[0] || c->max_batch < 2 * NHW_MAX_SUB) ? 1 : want_subs;
	if (subs > 1 && m < 2 * subs) subs = m >= 2 ? 2 : 1;
	p.slot = c->max_batch / subs;
	if (m > subs * p.slot) m = subs * p.slot;      // max_batch not a multiple of subs: the rest goes to the next wave
	p.count = m;
	p.subs = subs;
	p.streams = want_streams < subs ? want_streams : subs;
	for (int k = 0; k <= subs; k++) p.first[k] = (int)((long long)m * k / subs);   // consecutive differences <= ceil(m / subs) <= slot
	return p;
}

// A shallow copy of the context whose stream is lane l's and whose workspace pointers start at image `slot0`.
static nhw_ctx lane_view(const nhw_ctx *c, int l, int slot0, int sub = -1)
{
	nhw_ctx v = *c;
	const size_t s = (size_t)slot0;
	if (sub < 0) sub = l;
	v.stream = c->lanes[l];
	v.launches = 0;
	v.rgb += s * NHW_RGB_BYTES;
	v.y_jpeg += s * NHW_Y_SLOT; v.y_proc += s * NHW_Y_SLOT; v.y_aux += s * NHW_Y_SLOT; v.y_aux2 += s * NHW_Y_SLOT;
	v.y_ll1 += s * NHW_C_SLOT; v.y_ll2save += s * NHW_C_SLOT;
	v.c_u8 += s * 2 * NHW_CPLANE;
	v.c_jpeg += s * 2 * NHW_C_SLOT; v.c_proc += s * 2 * NHW_C_SLOT; v.c_aux += s * 2 * NHW_C_SLOT;
	v.c_ll1 += s * 2 * NHW_Q_SLOT; v.c_ll2save += s * 2 * NHW_Q_SLOT;
	v.rowmap += s * 512; v.rowcarry += s * 512;
	v.enc_bytes += s * (size_t)ENC_BYTES_SLOT; v.enc_hdr += s;
	v.dec_bytes += s * (size_t)NHW_DEC_BYTES_SLOT;
	v.out_dev += s * (size_t)NHW_MAX_STREAM_BYTES; v.pack_dev += s * (size_t)NHW_MAX_STREAM_BYTES;
	v.len_dev += s; v.status_dev += s; v.offs_dev += s + sub; v.offs_host += s + sub; v.status_host += s;
	v.dec_yuv += s * (size_t)NHW_RGB_BYTES;
	v.dec_desc_dev = static_cast<DecDesc *>(c->dec_desc_dev) + s;
	v.dec_desc_host = static_cast<DecDesc *>(c->dec_desc_host) + s;
	return v;
}

static void lanes_fork(nhw_ctx *c, int lanes)
{
	if (lanes <= 1) return;
	cudaEventRecord(c->ev_fork, c->lanes[0]);
	for (int l = 1; l < lanes; l++) cudaStreamWaitEvent(c->lanes[l], c->ev_fork, 0);
}
static void lanes_join(nhw_ctx *c, int lanes)
{
	for (int l = 1; l < lanes; l++) {
		cudaEventRecord(c->ev_join[l], c->lanes[l]);
		cudaStreamWaitEvent(c->lanes[0], c->ev_join[l], 0);
	}
}
static void lane_done(nhw_ctx *c, nhw_ctx &v)
{
	c->launches += v.launches;
	c->dbg_seen = v.dbg_seen;
	c->dbg_stopped = v.dbg_stopped;
}

static bool quality_supported(int q) { return q >= 1 && q <= 23; }   // q0 is accepted by the reference CLI but its tables are undefined

void *nhw_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (!check(cudaMallocHost(&p, bytes ? bytes : 1), "cudaMallocHost")) return nullptr;
	return p;
}

void nhw_host_free(void *p)
{
	if (p) cudaFreeHost(p);
}

int nhw_stage_colorspace_device(nhw_ctx *c, const uint8_t *rgb_dev, int n, int quality, int pre,
                                int16_t *y, uint8_t *u, uint8_t *v)
{
	if (!c || !rgb_dev || n <= 0) return NHW_ERR_ARG;
	if (quality < 1 || quality > 23) return NHW_ERR_QUALITY;
	cudaSetDevice(c->device);
	for (int i0 = 0; i0 < n; i0 += c->max_batch) {
		int m = n - i0 < c->max_batch ? n - i0 : c->max_batch;
		int16_t *yy = y ? y + (size_t)i0 * NHW_YPLANE : c->y_jpeg + NHW_GUARD_S;
		size_t ys = y ? (size_t)NHW_YPLANE : (size_t)NHW_Y_SLOT;
		nhw::colorspace(c, rgb_dev + (size_t)i0 * NHW_RGB_BYTES, m, quality, yy, ys,
		                u ? u + (size_t)i0 * NHW_CPLANE : nullptr, v ? v + (size_t)i0 * NHW_CPLANE : nullptr,
		                (size_t)NHW_CPLANE);
		if (pre && quality < 22) nhw::pre_processing(c, m, quality, yy, ys);
	}
	return finish(c, "nhw_stage_colorspace_device");
}

int nhw_stage_frontend_device(nhw_ctx *c, const uint8_t *rgb_dev, int n, int quality,
                              int16_t *y_proc, int16_t *y_ll1, int16_t *c_proc, int16_t *c_ll1)
{
	if (!c || !rgb_dev || n <= 0) return NHW_ERR_ARG;
	if (quality < 1 || quality > 23) return NHW_ERR_QUALITY;
	cudaSetDevice(c->device);
	const size_t YS = NHW_Y_SLOT, CS = NHW_C_SLOT, QS = NHW_Q_SLOT;
	for (int i0 = 0; i0 < n; i0 += c->max_batch) {
		int m = n - i0 < c->max_batch ? n - i0 : c->max_batch;
		// with caller buffers the planes are dense; otherwise they land in the context workspace
		int16_t *yp = y_proc ? y_proc + (size_t)i0 * NHW_YPLANE : c->y_proc + NHW_GUARD_S;
		int16_t *yl = y_ll1 ? y_ll1 + (size_t)i0 * NHW_CPLANE : c->y_ll1 + NHW_GUARD_S;
		int16_t *cp = c_proc ? c_proc + (size_t)i0 * 2 * NHW_CPLANE : c->c_proc + NHW_GUARD_S;
		int16_t *cl = c_ll1 ? c_ll1 + (size_t)i0 * 2 * 16384 : c->c_ll1 + NHW_GUARD_S;
		nhw::front_fused(c, rgb_dev + (size_t)i0 * NHW_RGB_BYTES, m, quality, yp, y_proc ? (size_t)NHW_YPLANE : YS, yl,
		                 y_ll1 ? (size_t)NHW_CPLANE : CS, c->c_u8, cp, c_proc ? (size_t)NHW_CPLANE : CS, cl,
		                 c_ll1 ? (size_t)16384 : QS, c->y_aux2 + NHW_GUARD_S, YS);
	}
	return finish(c, "nhw_stage_frontend_device");
}

int nhw_synth_batch_device(nhw_ctx *c, uint8_t *rgb_dev, int n, uint32_t seed0, int kind, const int16_t *sin_lut_dev)
{
	if (!c || !rgb_dev || n <= 0 || (kind != 1 && !sin_lut_dev)) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	nhw::synth(c, rgb_dev, n, seed0, kind, sin_lut_dev);
	return finish(c, "nhw_synth_batch_device");
}

int nhw_encode_batch_device(nhw_ctx *c, const uint8_t *rgb_dev, int n, int quality,
                            uint8_t *out_dev, uint32_t *len_dev, int32_t *status_dev)
{
	if (!c || !rgb_dev || !out_dev || n <= 0) return NHW_ERR_ARG;
	if (!quality_supported(quality)) { nhw::set_error("quality %d outside 1..23", quality); return NHW_ERR_QUALITY; }
	cudaSetDevice(c->device);
	c->dbg_seen = c->dbg_stopped = 0;
	for (int i0 = 0, step = 0; i0 < n; i0 += step) {
		// device-resident input: the kernels of one chunk already fill the GPU and share L2 better on one stream
		const int dl = c->tune.lanes_device;
		const LanePlan p = plan_lanes(c, n - i0 < c->max_batch ? n - i0 : c->max_batch, dl, dl);
		step = p.count;
		lanes_fork(c, p.streams);
		for (int l = 0; l < p.subs; l++) {
			const int a = i0 + p.first[l], cnt = p.first[l + 1] - p.first[l];
			if (cnt <= 0) continue;
			nhw_ctx v = lane_view(c, l, l * p.slot);
			v.chroma_side = (p.subs == 1 && !c->profile && !c->dbg_label[0] && c->tune.chroma_stream) ? 1 : 0;
			nhw::encode_chunk(&v, rgb_dev + (size_t)a * NHW_RGB_BYTES, cnt, quality, out_dev + (size_t)a * NHW_MAX_STREAM_BYTES,
			                  len_dev ? len_dev + a : nullptr, status_dev ? status_dev + a : nullptr);
			lane_done(c, v);
		}
		lanes_join(c, p.streams);
	}
	return finish(c, "nhw_encode_batch_device");
}

int nhw_pack_batch_device(nhw_ctx *c, const uint8_t *slots_dev, const uint32_t *len_dev, int n, uint64_t *offs_dev, uint8_t *dense_dev)
{
	if (!c || !slots_dev || !len_dev || !offs_dev || !dense_dev || n <= 0) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	nhw::pack_streams_to(c, slots_dev, len_dev, n, offs_dev, dense_dev);
	return finish(c, "nhw_pack_batch_device");
}

int nhw_digest_batch_device(nhw_ctx *c, const uint8_t *data_dev, size_t stride, const uint32_t *len_dev, uint32_t fixed_len, int n,
                            uint64_t *digest_dev)
{
	if (!c || !data_dev || !digest_dev || n <= 0) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	nhw::digest(c, data_dev, stride, len_dev, fixed_len, n, digest_dev);
	return finish(c, "nhw_digest_batch_device");
}

// ---- small host -> device transfers that must not queue behind another context's bulk uploads ----------------------
// Copies of one direction are served in issue order by the copy engine: a decode call's 13 MB of stream bytes would wait
// behind the gigabytes of pixel uploads an encode call on another context has already queued, and the two calls would
// run back to back instead of side by side.  Pinned (mapped) host memory is fetched by a kernel instead -- SM loads over
// PCIe do not pass through the copy queue; pageable memory keeps the cudaMemcpyAsync path.
__global__ void __launch_bounds__(256) k_fetch(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, size_t bytes)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, T = (size_t)gridDim.x * blockDim.x;
	if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
		const size_t n16 = bytes >> 4;
		const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
		uint4 *d4 = reinterpret_cast<uint4 *>(dst);
		size_t i = t;
		for (; i + 3 * T < n16; i += 4 * T) {   // four loads in flight per thread
			const uint4 a = s4[i], b = s4[i + T], c = s4[i + 2 * T], d = s4[i + 3 * T];
			d4[i] = a; d4[i + T] = b; d4[i + 2 * T] = c; d4[i + 3 * T] = d;
		}
		for (; i < n16; i += T) d4[i] = s4[i];
		for (size_t j = (n16 << 4) + t; j < bytes; j += T) dst[j] = src[j];
	} else {
		for (size_t j = t; j < bytes; j += T) dst[j] = src[j];
	}
}

// device-visible alias of a host pointer if the memory is pinned and mapped, else NULL
static const uint8_t *mapped_alias(const void *host)
{
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
	return static_cast<const uint8_t *>(at.devicePointer);
}

static bool fetch_to_device(nhw_ctx *w, void *dst, const void *host, const uint8_t *alias, size_t bytes, const char *what)
{
	if (!bytes) return true;
	if (!alias) return check(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, w->stream), what);
	const int blocks = bytes >= (8u << 20) ? 592 : bytes >= (64u << 10) ? 64 : 1;
	k_fetch<<<blocks, 256, 0, w->stream>>>(static_cast<uint8_t *>(dst), alias, bytes);
	w->launches++;
	return check(cudaGetLastError(), what);
}

int nhw_encode_batch(nhw_ctx *c, const uint8_t *rgb, int n, int quality,
                     uint8_t *out, size_t out_cap, uint64_t *offsets, int32_t *status)
{
	if (!c || !rgb || !out || !offsets || n <= 0) return NHW_ERR_ARG;
	if (!quality_supported(quality)) { nhw::set_error("quality %d outside 1..23", quality); return NHW_ERR_QUALITY; }
	cudaSetDevice(c->device);
	// Waves of sub-chunks dealt round robin to a few streams: a sub-chunk uploads its pixels, encodes and packs
	// on its stream, so uploads overlap the kernels of the other streams; finished sub-chunks come back on the
	// copy stream, in image order.
	uint64_t pos = 0;
	offsets[0] = 0;
	const uint8_t *out_alias = c->tune.fetch_kernel ? mapped_alias(out) : nullptr;
	for (int i0 = 0, step = 0; i0 < n; i0 += step) {
		const LanePlan p = plan_lanes(c, n - i0 < c->max_batch ? n - i0 : c->max_batch, c->tune.subs_encode,
		                              c->tune.lanes_encode);
		step = p.count;
		nhw_ctx v[NHW_MAX_SUB];
		lanes_fork(c, p.streams);
		// every sub-chunk has its own pixel slice: all uploads are queued back to back on their own stream, so the
		// copy engine never waits for the kernels of an earlier sub-chunk that shares a stream with a later one
		if (p.streams <= 1) cudaEventRecord(c->ev_fork, c->lanes[0]);
		cudaStreamWaitEvent(c->up_stream, c->ev_fork, 0);
		for (int k = 0; k < p.subs; k++) {
			const int a = i0 + p.first[k], cnt = p.first[k + 1] - p.first[k];
			v[k] = lane_view(c, k % p.streams, k * p.slot, k);
			if (cnt <= 0) continue;
			if (!check(cudaMemcpyAsync(v[k].rgb, rgb + (size_t)a * NHW_RGB_BYTES, (size_t)cnt * NHW_RGB_BYTES,
			                           cudaMemcpyHostToDevice, c->up_stream), "H2D pixels")) return NHW_ERR_CUDA;
			cudaEventRecord(c->ev_up[k], c->up_stream);
		}
		for (int k = 0; k < p.subs; k++) {
			const int cnt = p.first[k + 1] - p.first[k];
			if (cnt <= 0) continue;
			cudaStreamWaitEvent(v[k].stream, c->ev_up[k], 0);
			nhw::encode_chunk(&v[k], v[k].rgb, cnt, quality, v[k].out_dev, v[k].len_dev, v[k].status_dev);
			nhw::pack_streams(&v[k], cnt);
			// the sub-chunk's offsets and status words go back by SM stores into the (mapped) staging arrays, not through the
			// copy engine: behind another context's bulk downloads these few kilobytes would hold this lane up for milliseconds
			if (c->tune.fetch_kernel) {
				k_fetch<<<1, 256, 0, v[k].stream>>>(reinterpret_cast<uint8_t *>(v[k].offs_host), reinterpret_cast<const uint8_t *>(v[k].offs_dev),
				                                    (size_t)(cnt + 1) * sizeof(uint64_t));
				k_fetch<<<1, 256, 0, v[k].stream>>>(reinterpret_cast<uint8_t *>(v[k].status_host), reinterpret_cast<const uint8_t *>(v[k].status_dev),
				                                    (size_t)cnt * sizeof(int32_t));
				v[k].launches += 2;
			} else {
				cudaMemcpyAsync(v[k].offs_host, v[k].offs_dev, (size_t)(cnt + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, v[k].stream);
				cudaMemcpyAsync(v[k].status_host, v[k].status_dev, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, v[k].stream);
			}
			cudaEventRecord(c->ev_sub[k], v[k].stream);
			lane_done(c, v[k]);
		}
		for (int k = 0; k < p.subs; k++) {
			const int a = i0 + p.first[k], cnt = p.first[k + 1] - p.first[k];
			if (cnt <= 0) continue;
			if (!check(cudaGetLastError(), "nhw_encode_batch") || !check(cudaEventSynchronize(c->ev_sub[k]), "nhw_encode_batch")) return NHW_ERR_CUDA;
			const uint64_t total = v[k].offs_host[cnt];
			if (pos + total > out_cap) { nhw::set_error("output buffer too small"); return NHW_ERR_ARG; }
			if (total && out_alias) {   // pinned output buffer: SM stores again (see k_fetch), a few megabytes per sub-chunk
				k_fetch<<<total >= (1u << 20) ? 128 : 8, 256, 0, c->copy_stream>>>(const_cast<uint8_t *>(out_alias) + pos, v[k].pack_dev, total);
				c->launches++;
				if (!check(cudaGetLastError(), "D2H streams")) return NHW_ERR_CUDA;
			} else if (total && !check(cudaMemcpyAsync(out + pos, v[k].pack_dev, total, cudaMemcpyDeviceToHost, c->copy_stream), "D2H streams")) return NHW_ERR_CUDA;
			for (int i = 0; i < cnt; i++) {
				offsets[a + i + 1] = pos + v[k].offs_host[i + 1];
				if (status) status[a + i] = v[k].status_host[i];
			}
			pos += total;
		}
		if (!check(cudaStreamSynchronize(c->copy_stream), "nhw_encode_batch")) return NHW_ERR_CUDA;
		for (int l = 0; l < p.streams; l++)
			if (!check(cudaStreamSynchronize(c->lanes[l]), "nhw_encode_batch")) return NHW_ERR_CUDA;
	}
	return NHW_OK;
}

static int decode_batch_impl(nhw_ctx *c, const uint8_t *in, const uint64_t *offsets, int n, uint8_t *rgb, uint8_t *yuv,
                             int32_t *quality, int32_t *status)
{
	if (!c || !in || !offsets || (!rgb && !yuv) || n <= 0) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	c->dbg_seen = c->dbg_stopped = 0;
	const uint8_t *in_alias = c->tune.fetch_kernel ? mapped_alias(in) : nullptr;
	// the context's own staging arrays come from cudaMallocHost: mapped, and (unified addressing) at the same address
	auto ctx_alias = [&](const void *p) { return in_alias ? static_cast<const uint8_t *>(p) : nullptr; };
	for (int i0 = 0, step = 0; i0 < n; i0 += step) {
		const LanePlan p = plan_lanes(c, n - i0 < c->max_batch ? n - i0 : c->max_batch, c->tune.subs_decode,
		                              c->tune.lanes_decode);
		step = p.count;
		nhw_ctx v[NHW_MAX_SUB];
		lanes_fork(c, p.streams);
		for (int l = 0; l < p.subs; l++) {
			const int a = i0 + p.first[l], cnt = p.first[l + 1] - p.first[l];
			v[l] = lane_view(c, l % p.streams, l * p.slot, l);
			if (cnt <= 0) continue;
			nhw_ctx &w = v[l];
			DecDesc *desc = static_cast<DecDesc *>(w.dec_desc_host);
			const uint64_t base = offsets[a], total = offsets[a + cnt] - base;
			const uint64_t mis = in_alias ? ((uintptr_t)(in + base) & 15) : 0;
			if (total + mis + 64 > (uint64_t)p.slot * NHW_MAX_STREAM_BYTES) { nhw::set_error("input chunk too large"); return NHW_ERR_ARG; }
			bool any_lowq = false, any_hq = false;   // q <= 16 and q >= 22 streams each take one extra kernel
			for (int i = 0; i < cnt; i++) {
				const uint64_t o = offsets[a + i] - base, len = offsets[a + i + 1] - offsets[a + i];
				w.offs_host[i] = o + mis;
				w.status_host[i] = nhw_parse_header(in + base + o, (size_t)len, &desc[i]);
				if (quality) quality[a + i] = desc[i].quality;
				any_lowq |= w.status_host[i] == 0 && desc[i].quality <= 16;
				any_hq |= w.status_host[i] == 0 && desc[i].quality >= 22;
			}
			// (mis: the chunk lands at the same offset modulo 16 as it has in the caller's buffer, so the fetch kernel moves
			// aligned 16-byte words; the per-stream offsets carry the shift)
			bool ok = fetch_to_device(&w, w.pack_dev + mis, in + base, in_alias ? in_alias + base : nullptr, total, "H2D streams");
			// the bit reader may look a few words past the last code: keep that tail defined
			ok = ok && check(cudaMemsetAsync(w.pack_dev + mis + total, 0, 64, w.stream), "memset tail");
			ok = ok && fetch_to_device(&w, w.dec_desc_dev, desc, ctx_alias(desc), (size_t)cnt * sizeof(DecDesc), "H2D desc");
			ok = ok && fetch_to_device(&w, w.offs_dev, w.offs_host, ctx_alias(w.offs_host), (size_t)cnt * sizeof(uint64_t), "H2D offs");
			ok = ok && fetch_to_device(&w, w.status_dev, w.status_host, ctx_alias(w.status_host), (size_t)cnt * sizeof(int32_t), "H2D status");
			if (!ok) return NHW_ERR_CUDA;
			w.chroma_side = 0;
			nhw::decode_chunk(&w, w.pack_dev, w.offs_dev, static_cast<const DecDesc *>(w.dec_desc_dev), w.status_dev, cnt, w.rgb, any_lowq, any_hq, yuv != nullptr);
			if (rgb) cudaMemcpyAsync(rgb + (size_t)a * NHW_RGB_BYTES, w.rgb, (size_t)cnt * NHW_RGB_BYTES, cudaMemcpyDeviceToHost, w.stream);
			if (yuv) cudaMemcpyAsync(yuv + (size_t)a * NHW_RGB_BYTES, w.dec_yuv, (size_t)cnt * NHW_RGB_BYTES, cudaMemcpyDeviceToHost, w.stream);
			cudaMemcpyAsync(w.status_host, w.status_dev, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, w.stream);
			lane_done(c, w);
		}
		for (int l = 0; l < p.streams; l++)
			if (!check(cudaGetLastError(), "nhw_decode_batch") || !check(cudaStreamSynchronize(c->lanes[l]), "nhw_decode_batch")) return NHW_ERR_CUDA;
		for (int l = 0; l < p.subs; l++) {
			const int a = i0 + p.first[l], cnt = p.first[l + 1] - p.first[l];
			for (int i = 0; status && i < cnt; i++) status[a + i] = v[l].status_host[i];
		}
	}
	return NHW_OK;
}

static int decode_device_impl(nhw_ctx *c, const uint8_t *in_dev, size_t stride, const uint32_t *len_dev, const uint64_t *offs_dev,
                              int n, uint8_t *rgb_dev, int32_t *status_dev)
{
	if (!c || !in_dev || !rgb_dev || n <= 0 || (!len_dev && !offs_dev)) return NHW_ERR_ARG;
	cudaSetDevice(c->device);
	c->dbg_seen = c->dbg_stopped = 0;
	for (int i0 = 0; i0 < n; i0 += c->max_batch) {
		const int m = n - i0 < c->max_batch ? n - i0 : c->max_batch;
		nhw::decode_chunk_device(c, offs_dev ? in_dev : in_dev + (size_t)i0 * stride, stride, len_dev ? len_dev + i0 : nullptr,
		                         offs_dev ? offs_dev + i0 : nullptr, m, rgb_dev + (size_t)i0 * NHW_RGB_BYTES,
		                         status_dev ? status_dev + i0 : nullptr);
	}
	return finish(c, "nhw_decode_batch_device");
}

int nhw_decode_batch_device(nhw_ctx *c, const uint8_t *in_dev, size_t stride, const uint32_t *len_dev, int n, uint8_t *rgb_dev,
                            int32_t *status_dev)
{
	if (!len_dev || stride == 0) return NHW_ERR_ARG;
	return decode_device_impl(c, in_dev, stride, len_dev, nullptr, n, rgb_dev, status_dev);
}

int nhw_decode_batch_packed_device(nhw_ctx *c, const uint8_t *in_dev, const uint64_t *offs_dev, int n, uint8_t *rgb_dev,
                                   int32_t *status_dev)
{
	if (!offs_dev) return NHW_ERR_ARG;
	return decode_device_impl(c, in_dev, 0, nullptr, offs_dev, n, rgb_dev, status_dev);
}

int nhw_decode_batch(nhw_ctx *c, const uint8_t *in, const uint64_t *offsets, int n, uint8_t *rgb, int32_t *status)
{
	if (!rgb) return NHW_ERR_ARG;
	return decode_batch_impl(c, in, offsets, n, rgb, nullptr, nullptr, status);
}

int nhw_decode_batch_planes(nhw_ctx *c, const uint8_t *in, const uint64_t *offsets, int n, uint8_t *yuv, int32_t *quality,
                            int32_t *status)
{
	if (!yuv) return NHW_ERR_ARG;
	return decode_batch_impl(c, in, offsets, n, nullptr, yuv, quality, status);
}

}  // extern "C"
