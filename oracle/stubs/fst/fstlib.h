/*
 * TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for OpenFst's <fst/fstlib.h>,
 * which ASR-CRaFT includes (CRF/src/nodes/CRF_StateNode.h:13, CRF/src/decoders/CRF_ViterbiDecoder.h:13)
 * but does not vendor.  It provides just enough of VectorFst<StdArc> for the token-passing
 * Viterbi decoders to build their free-phone LM and their linear best-path output
 * (CRF/src/decoders/CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:1270-1348, 2204-2354).
 * No shortest-path / composition algorithm is implemented: the decoder's own best path is
 * already a single linear chain, so Compose() hands back its first argument and Connect()
 * is the identity.  Written from the call sites, not from OpenFst sources.
 */
#ifndef ORACLE_STUB_FSTLIB_H
#define ORACLE_STUB_FSTLIB_H

#include <stdint.h>
#include <sys/types.h>
#include <limits>
#include <vector>
#include <list>
#include <string>

typedef int64_t int64;
typedef uint64_t uint64;
typedef int32_t int32;
typedef uint32_t uint32;

namespace fst {

const int kNoStateId = -1;
const int kNoLabel = -1;

class TropicalWeight {
	float v_;
public:
	TropicalWeight() : v_(std::numeric_limits<float>::infinity()) {}
	TropicalWeight(float v) : v_(v) {}
	TropicalWeight(double v) : v_((float)v) {}
	TropicalWeight(int v) : v_((float)v) {}
	float Value() const { return v_; }
	static TropicalWeight Zero() { return TropicalWeight(std::numeric_limits<float>::infinity()); }
	static TropicalWeight One() { return TropicalWeight(0.0f); }
	bool operator==(const TropicalWeight& o) const { return v_ == o.v_; }
	bool operator!=(const TropicalWeight& o) const { return v_ != o.v_; }
};

struct StdArc {
	typedef int StateId;
	typedef int Label;
	typedef TropicalWeight Weight;
	Label ilabel;
	Label olabel;
	Weight weight;
	StateId nextstate;
	StdArc() : ilabel(0), olabel(0), weight(), nextstate(kNoStateId) {}
	StdArc(Label i, Label o, Weight w, StateId n) : ilabel(i), olabel(o), weight(w), nextstate(n) {}
};
typedef StdArc LogArc;

template <class A> class Fst {
public:
	typedef A Arc;
	std::vector<std::vector<A> > arcs_;
	std::vector<typename A::Weight> finals_;
	int start_;
	Fst() : start_(kNoStateId) {}
	virtual ~Fst() {}
};

template <class A> class VectorFst : public Fst<A> {
public:
	typedef typename A::Weight Weight;
	int AddState() {
		this->arcs_.push_back(std::vector<A>());
		this->finals_.push_back(Weight::Zero());
		return (int)this->arcs_.size() - 1;
	}
	void AddArc(int s, const A& a) { this->arcs_[s].push_back(a); }
	void SetStart(int s) { this->start_ = s; }
	int Start() const { return this->start_; }
	void SetFinal(int s, Weight w) { this->finals_[s] = w; }
	Weight Final(int s) const { return this->finals_[s]; }
	int NumStates() const { return (int)this->arcs_.size(); }
	size_t NumArcs(int s) const { return this->arcs_[s].size(); }
	void DeleteStates(const std::vector<int>&) {}
	void DeleteStates() { this->arcs_.clear(); this->finals_.clear(); this->start_ = kNoStateId; }
	bool Write(const std::string&) const { return true; }
};

template <class F> class ArcIterator {
	const std::vector<typename F::Arc>* v_;
	size_t i_;
public:
	ArcIterator(const F& f, int s) : v_(&f.arcs_[s]), i_(0) {}
	bool Done() const { return i_ >= v_->size(); }
	void Next() { ++i_; }
	const typename F::Arc& Value() const { return (*v_)[i_]; }
};

template <class F> class MutableArcIterator {
	std::vector<typename F::Arc>* v_;
	size_t i_;
public:
	MutableArcIterator(F* f, int s) : v_(&f->arcs_[s]), i_(0) {}
	bool Done() const { return i_ >= v_->size(); }
	void Next() { ++i_; }
	const typename F::Arc& Value() const { return (*v_)[i_]; }
	void SetValue(const typename F::Arc& a) { (*v_)[i_] = a; }
};

template <class F> class StateIterator {
	int n_, i_;
public:
	StateIterator(const F& f) : n_((int)f.arcs_.size()), i_(0) {}
	bool Done() const { return i_ >= n_; }
	void Next() { ++i_; }
	int Value() const { return i_; }
};

template <class F> void Connect(F*) {}
template <class A, class B, class C> void Compose(const A& a, const B&, C* c) {
	c->arcs_ = a.arcs_; c->finals_ = a.finals_; c->start_ = a.start_;
}

}  // namespace fst

#endif
