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

/* CRF_FeatureStream::join() (CRF/src/io/CRF_FeatureStream.cpp:172-184) names this class; the
 * oracle never joins streams, so the methods abort if reached. */
class QN_InFtrStream_JoinFtrs : public QN_InFtrStream {
public:
	QN_InFtrStream_JoinFtrs(int, const char*, QN_InFtrStream&, QN_InFtrStream&) {}
	size_t num_ftrs() { abort(); }
	QN_SegID nextseg() { abort(); }
	size_t read_ftrs(size_t, float*) { abort(); }
	int rewind() { abort(); }
	size_t num_segs() { abort(); }
	size_t num_frames(size_t = QN_ALL) { abort(); }
	int get_pos(size_t*, size_t*) { abort(); }
	QN_SegID set_pos(size_t, size_t) { abort(); }
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
