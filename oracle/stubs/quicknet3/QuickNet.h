/*
 * TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for the ICSI QuickNet3
 * header, which ASR-CRaFT includes from CRF/src/CRF.h:30 but does not vendor.  Only the
 * surface used by the CRF lattice hot path is declared: integer typedefs, stream ids,
 * the abstract feature/label stream interfaces consumed by CRF_FeatureStream
 * (CRF/src/io/CRF_FeatureStream.h:39-45) and by the SeqMultiWindow streams
 * (CRF/src/io/CRF_InFtrStream_SeqMultiWindow.h:13-29), a logger and one vector copy.
 * Written from the call sites, not from QuickNet sources.
 */
#ifndef ORACLE_STUB_QUICKNET_H
#define ORACLE_STUB_QUICKNET_H

#include <stdint.h>
#include <stddef.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t QNUInt32;
typedef int32_t QNInt32;
typedef long QN_SegID;

enum { QN_SEGID_BAD = -1 };
enum { QN_OK = 0, QN_BAD = -1 };
#define QN_ALL ((size_t)ULONG_MAX)
#define QN_SIZET_BAD ((size_t)ULONG_MAX)
#ifndef QN_UINT32_MAX
#define QN_UINT32_MAX (0xffffffff)
#endif
#define QN_OUTPUT(...) do { fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } while (0)

class QN_InFtrStream {
public:
	virtual ~QN_InFtrStream() {}
	virtual size_t num_ftrs() = 0;
	virtual QN_SegID nextseg() = 0;
	virtual size_t read_ftrs(size_t cnt, float* ftrs) = 0;
	virtual int rewind() = 0;
	virtual size_t num_segs() = 0;
	virtual size_t num_frames(size_t segno = QN_ALL) = 0;
	virtual int get_pos(size_t* segno, size_t* frameno) = 0;
	virtual QN_SegID set_pos(size_t segno, size_t frameno) = 0;
};

class QN_InLabStream {
public:
	virtual ~QN_InLabStream() {}
	virtual size_t num_labs() = 0;
	virtual QN_SegID nextseg() = 0;
	virtual size_t read_labs(size_t cnt, QNUInt32* labs) = 0;
	virtual int rewind() = 0;
	virtual size_t num_segs() = 0;
	virtual size_t num_frames(size_t segno = QN_ALL) = 0;
	virtual int get_pos(size_t* segno, size_t* frameno) = 0;
	virtual QN_SegID set_pos(size_t segno, size_t frameno) = 0;
};

/* CRF_FeatureStream::join() (CRF/src/io/CRF_FeatureStream.cpp:172-184) names this QuickNet3 class (version unpinned by the
 * reference, sources absent).  Its published behaviour, restated from the call sites: the two streams advance together segment by
 * segment and frame by frame, and every frame read is the first stream's feature vector followed by the second's. */
class QN_InFtrStream_JoinFtrs : public QN_InFtrStream {
	QN_InFtrStream& a_; QN_InFtrStream& b_;
	float* ta_; float* tb_; size_t cap_;
public:
	QN_InFtrStream_JoinFtrs(int, const char*, QN_InFtrStream& a, QN_InFtrStream& b) : a_(a), b_(b), ta_(NULL), tb_(NULL), cap_(0) {}
	~QN_InFtrStream_JoinFtrs() { free(ta_); free(tb_); }
	size_t num_ftrs() { return a_.num_ftrs() + b_.num_ftrs(); }
	QN_SegID nextseg() {
		QN_SegID x = a_.nextseg(), y = b_.nextseg();
		if (x != y) { fputs("QN_InFtrStream_JoinFtrs: the joined streams disagree on the segment id\n", stderr); abort(); }
		return x;
	}
	size_t read_ftrs(size_t cnt, float* ftrs) {
		const size_t wa = a_.num_ftrs(), wb = b_.num_ftrs();
		if (cnt > cap_) { free(ta_); free(tb_); cap_ = cnt; ta_ = (float*)malloc(cap_ * wa * sizeof(float)); tb_ = (float*)malloc(cap_ * wb * sizeof(float)); }
		const size_t na = a_.read_ftrs(cnt, ta_), nb = b_.read_ftrs(cnt, tb_);
		if (na != nb) { fputs("QN_InFtrStream_JoinFtrs: the joined streams disagree on the frame count\n", stderr); abort(); }
		for (size_t i = 0; i < na; i++) {
			memcpy(ftrs + i * (wa + wb), ta_ + i * wa, wa * sizeof(float));
			memcpy(ftrs + i * (wa + wb) + wa, tb_ + i * wb, wb * sizeof(float));
		}
		return na;
	}
	int rewind() { int x = a_.rewind(); int y = b_.rewind(); return (x == QN_OK && y == QN_OK) ? QN_OK : QN_BAD; }
	size_t num_segs() { return a_.num_segs(); }
	size_t num_frames(size_t segno = QN_ALL) { return a_.num_frames(segno); }
	int get_pos(size_t* s, size_t* f) { return a_.get_pos(s, f); }
	QN_SegID set_pos(size_t s, size_t f) { b_.set_pos(s, f); return a_.set_pos(s, f); }
};

enum { QN_LOG_PER_RUN = 1, QN_LOG_PER_EPOCH, QN_LOG_PER_SENT, QN_LOG_PER_BUNCH };

class QN_ClassLogger {
public:
	QN_ClassLogger(int, const char*, const char*) {}
	void log(int, const char*, ...) {}
	void warning(const char*, ...) {}
	void error(const char* fmt, ...) {
		va_list ap; va_start(ap, fmt);
		fputs("QN_ClassLogger::error: ", stderr); vfprintf(stderr, fmt, ap); fputc('\n', stderr);
		va_end(ap); abort();
	}
};

inline void qn_copy_vf_vf(size_t n, const float* from, float* to) { memmove(to, from, n * sizeof(float)); }

#endif
