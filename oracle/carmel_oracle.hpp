// carmel_oracle.hpp -- CPU ORACLE (test infrastructure, NOT the product).
//
// A plain C++17, Boost-free restatement of the reference algorithm for carmel's training hot
// path (graehl/carmel), used only by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.
// Nothing under carmel_b200/ may include, link or execute this file.
//
// Parity status: PINNED for EM (single WFST and cascades) against the reference's own golden
// run log carmel/carmel-tutorial/commands.trace (epron-jpron 5 iterations, cipher cascade 22
// iterations, tagging cascade 9 iterations; see tests/test_oracle_golden.py).  Gibbs sampled
// derivations are "parity unpinned" by the reference (its log is RNG dependent, seed not
// recorded): they are pinned only to this restatement with injected uniforms.
//
// Each function cites the reference file:line it follows (paths relative to the reference root).
// Data structures deliberately mirror the reference's class of structure (per-example adjacency
// lists, log-space fp64 Weight, recursive DFS trellis construction) so that timing this oracle
// is a fair stand-in for the reference binary, which cannot be built here (needs Boost).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------
// logweight<double>  (graehl/shared/weight.h:131-604; ops :737-830; config.h:41,149)
// ------------------------------------------------------------------------------------------
static const double ORC_INF = std::numeric_limits<double>::infinity();
static const double MUCH_BIGGER_LN = 36.;  // weight.h:103 (double)
static const double UNDERFLOW_LN = 82.;  // weight.h:112

struct W {
  double w;  // natural log of the represented non-negative real
  W() : w(-ORC_INF) {}  // weight.h:272-292: default = zero
  explicit W(double ln, bool) : w(ln) {}
  W(double real) { setReal(real); }  // weight.h:295-300
  static W ln(double l) { return W(l, false); }
  static W one() { return W(0., false); }
  static W zero() { return W(); }
  static W inf() { return W(ORC_INF, false); }
  void setReal(double f) { w = f > 0 ? std::log(f) : -ORC_INF; }
  void setZero() { w = -ORC_INF; }
  void setOne() { w = 0; }
  void setInfinity() { w = ORC_INF; }
  bool isZero() const { return !(w > -ORC_INF); }
  bool isPositive() const { return w > -ORC_INF; }
  bool isInfinity() const { return w == ORC_INF; }
  bool isOne() const { return w == 0; }
  double getReal() const { return std::exp(w); }
  double getLn() const { return w; }
  bool fitsInReal() const { return isZero() || (w < UNDERFLOW_LN && w > -UNDERFLOW_LN); }  // weight.h:266
  W pow(double n) const { return isZero() ? *this : W(w * n, false); }  // weight.h:442-447
  W root(double n) const { return isZero() ? *this : W(w / n, false); }  // weight.h:435-440
  W ppxper(double n = 1) const { return root(-n); }  // weight.h:309
  // weight.h:247-249
  W relative_perplexity_ratio(W const& o) const;
};
inline W operator*(W a, W b) { return W(a.w + b.w, false); }  // weight.h:737
inline W operator/(W a, W b) { return W(a.w - b.w, false); }  // weight.h:738
inline W operator+(W lhs, W rhs) {  // weight.h:765-801 (WEIGHT_CORRECT_ZERO, GRAEHL_USE_LOG1P)
  if (lhs.isZero()) return rhs;
  if (rhs.isZero()) return lhs;
  double diff = lhs.w - rhs.w;
  if (diff > MUCH_BIGGER_LN) return lhs;
  if (diff < -MUCH_BIGGER_LN) return rhs;
  if (diff < 0) return W(rhs.w + log1p(std::exp(diff)), false);
  return W(lhs.w + log1p(std::exp(-diff)), false);
}
inline W operator-(W lhs, W rhs) {  // weight.h:803-830
  if (rhs.isZero()) return lhs;
  double rdiff = rhs.w - lhs.w;
  if (rdiff >= 0) return W();
  if (rdiff < -MUCH_BIGGER_LN) return lhs;
  return W(lhs.w + log1p(-std::exp(rdiff)), false);
}
inline W& operator+=(W& a, W b) { return a = a + b; }
inline W& operator-=(W& a, W b) { return a = a - b; }
inline W& operator*=(W& a, W b) { return a = a * b; }
inline W& operator/=(W& a, W b) { return a = a / b; }
inline bool operator<(W a, W b) { return a.w < b.w; }
inline bool operator>(W a, W b) { return a.w > b.w; }
inline bool operator<=(W a, W b) { return a.w <= b.w; }
inline bool operator>=(W a, W b) { return a.w >= b.w; }
inline bool operator==(W a, W b) { return a.w == b.w; }
inline bool operator!=(W a, W b) { return a.w != b.w; }
inline W absdiff(W a, W b) { return a.w > b.w ? a - b : b - a; }  // weight.h:836-855
inline W W::relative_perplexity_ratio(W const& o) const { return (*this / o).root(std::fabs(w)); }

// weight.h:467-490 print (default: EXP base, SOMETIMES_LOG as set by carmel.cc setOutputFormat)
inline std::string fmt_double(double d, int prec) {
  std::ostringstream o;
  o.precision(prec);
  o << d;
  return o.str();
}
struct WeightFormat {
  enum { EXP, LN, LOG10 } base = EXP;
  enum { SOMETIMES, ALWAYS, NEVER } thresh = SOMETIMES;
};
inline std::string fmt_weight(W x, WeightFormat const& f = WeightFormat()) {
  if (x.isZero()) return "0";
  if ((f.thresh == WeightFormat::SOMETIMES && x.fitsInReal()) || f.thresh == WeightFormat::NEVER)
    return fmt_double(x.getReal(), 15);
  if (f.base == WeightFormat::LN) return fmt_double(x.w, 15) + "ln";
  if (f.base == WeightFormat::LOG10) return fmt_double(x.w / 2.30258509299404568402, 15) + "log";
  return "e^" + fmt_double(x.w, 15);
}
// weight.h:546-549,592-601 as_base(2) printed at the stream's default precision (6)
inline std::string fmt_base2(W x) { return "2^" + fmt_double(x.w / std::log(2.), 6); }

// weight.h:503-528 setStringPartial / setString
inline bool parse_weight(const char* b, W& out) {
  const char* end = b + std::strlen(b);
  char* e;
  if (b == end) return false;
  if (b + 1 < end && b[0] == 'e' && b[1] == '^') {
    out = W::ln(std::strtod(b + 2, &e));
    return e == end && e != b + 2;
  } else if (b + 2 < end && b[0] == '1' && b[1] == '0' && b[2] == '^') {
    out = W::ln(std::strtod(b + 3, &e) * 2.30258509299404568402);
    return e == end && e != b + 3;
  } else {
    double d = std::strtod(b, &e);
    if (e == b) return false;
    if (e[0] == 'l') {
      if (e[1] == 'n') {
        out = W::ln(d);
        return e + 2 == end;
      } else if (e[1] == 'o' && e[2] == 'g') {
        out = W::ln(d * 2.30258509299404568402);
        return e + 3 == end;
      }
      return false;
    }
    out = W(d);
    return e == end;
  }
}

// ------------------------------------------------------------------------------------------
// Alphabet / FSTArc / WFST  (graehl/shared/arc.h:28-70; carmel/src/fst.h:409,483-490; state.h)
// ------------------------------------------------------------------------------------------
static const unsigned NO_GROUP = 0xFFFFFFFFu;  // arc.h:43
static const unsigned LOCKED_GROUP = 0;  // arc.h:44
static const unsigned EPS = 0;  // arc.h:45

struct Alphabet {
  std::vector<std::string> names;
  std::unordered_map<std::string, unsigned> idx;
  Alphabet() {
    index_of("*e*");
    index_of("*w*");
  }
  unsigned index_of(std::string const& s) {
    auto it = idx.find(s);
    if (it != idx.end()) return it->second;
    unsigned i = (unsigned)names.size();
    names.push_back(s);
    idx.emplace(s, i);
    return i;
  }
  int find(std::string const& s) const {
    auto it = idx.find(s);
    return it == idx.end() ? -1 : (int)it->second;
  }
};

struct Arc {
  unsigned in, out, dest;
  W weight;
  unsigned group;
  bool isLocked() const { return group == LOCKED_GROUP; }
  bool isNormal() const { return group == NO_GROUP; }
  bool isTied() const { return !isLocked() && !isNormal(); }
};

enum NormGroupBy { CONDITIONAL, JOINT, NONE };  // fst.h:552-556
struct NormalizeMethod {  // fst.h:568-580 (scale = identity: digamma mode is out of scope)
  NormGroupBy group = CONDITIONAL;
  W add_count;  // --priors
};

struct WFST {
  typedef std::deque<Arc> Arcs;  // reader appends (state.h:209-231); compose pushes front (state.h:234)
  std::vector<Arcs> states;
  std::vector<std::string> stateNames;
  std::unordered_map<std::string, unsigned> stateIdx;
  bool named = true;
  unsigned final_state = 0;
  bool is_valid = false;
  std::shared_ptr<Alphabet> alph[2];

  WFST() {
    alph[0] = std::make_shared<Alphabet>();
    alph[1] = std::make_shared<Alphabet>();
  }
  bool valid() const { return is_valid; }
  unsigned numStates() const { return (unsigned)states.size(); }
  size_t numArcs() const {
    size_t n = 0;
    for (auto const& s : states) n += s.size();
    return n;
  }
  std::string stateName(unsigned i) const {
    if (named && i < stateNames.size()) return stateNames[i];
    return std::to_string(i);
  }

  // ---- text reader: carmel/src/wfstio.cc:95-150 (getString), :341-506 (readLegible) ----
  static bool getString(std::istream& in, std::string& out) {
    char c;
    out.clear();
    if (!(in >> c)) return false;
    switch (c) {
      case '"': {
        out.push_back(c);
        bool l = false;
        for (;;) {
          char s;
          if (!in.get(s)) return false;
          out.push_back(s);
          if (s == '"' && !l) break;
          if (s == '\\')
            l = !l;
          else
            l = false;
        }
        return true;
      }
      case '*': {
        out.push_back(c);
        for (;;) {
          char s;
          if (!in.get(s)) return false;
          if (s == '*') {
            out.push_back(s);
            break;
          }
          out.push_back((char)tolower(s));
        }
        return true;
      }
      case '(':
      case ')': return false;
      default: {
        out.push_back(c);
        char s;
        while (in.get(s)) {
          if (s == '\n' || s == '\t' || s == ' ') break;
          if (s == '!' || s == ')') {
            in.unget();
            break;
          }
          out.push_back(s);
        }
        if (!out.empty() && out.back() == '\r') out.pop_back();
        return true;
      }
    }
  }
  static void skip_comment(std::istream& in) {  // '%' to end of line, repeatedly
    for (;;) {
      char c;
      if (!(in >> c)) return;
      if (c == '%') {
        std::string dummy;
        std::getline(in, dummy);
      } else {
        in.unget();
        return;
      }
    }
  }
  unsigned getStateIndex(std::string const& buf) {  // wfstio.cc:313-336
    if (!named) {
      char* e;
      unsigned long st = strtol(buf.c_str(), &e, 10);
      if (!buf.empty() && *e != '\0') return ~0u;
      if (st >= states.size()) states.resize(st + 1);
      return (unsigned)st;
    }
    auto it = stateIdx.find(buf);
    if (it != stateIdx.end()) return it->second;
    unsigned i = (unsigned)stateNames.size();
    stateNames.push_back(buf);
    stateIdx.emplace(buf, i);
    states.emplace_back();
    return i;
  }
  bool read(std::istream& istr, bool alwaysNamed = true) {
    Alphabet &in = *alph[0], &out = *alph[1];
    std::string buf, buf2, finalName;
    named = true;
    is_valid = false;
    char c;
    skip_comment(istr);
    if (!getString(istr, buf)) return false;
    finalName = buf;
    if (!alwaysNamed) {
      named = false;
      for (char ch : finalName)
        if (!isdigit((unsigned char)ch)) {
          named = true;
          break;
        }
    }
    if (!named) final_state = getStateIndex(buf);
#define ORC_REQ(x) \
  do {             \
    if (!(x)) return false; \
  } while (0)
#define ORC_GETC ORC_REQ(istr >> c)
#define ORC_PEEKC \
  do {            \
    ORC_REQ(istr >> c); \
    istr.unget(); \
  } while (0)
    while (istr >> c) {
      // note: the reference calls skip_comment after reading '(' (wfstio.cc:376); comments between lines
      if (c == '%') {
        std::string dummy;
        std::getline(istr, dummy);
        continue;
      }
      ORC_REQ(c == '(');
      ORC_REQ(getString(istr, buf));
      unsigned src = getStateIndex(buf);
      ORC_REQ(~src);
      for (;;) {
        ORC_GETC;
        bool destparen = (c == '(');
        if (!destparen) istr.unget();
        if (c == ')') break;
        ORC_REQ(getString(istr, buf));
        unsigned dst = getStateIndex(buf);
        ORC_REQ(~dst);
        for (;;) {
          ORC_GETC;
          bool iowparen = (c == '(');
          if (!iowparen)
            istr.unget();
          else
            ORC_PEEKC;
          unsigned inL, outL;
          W weight = W::one();
          auto endiow = [&]() { return c == ')' || c == '!'; };
          if (endiow()) {
            inL = outL = EPS;
          } else {
            ORC_REQ(getString(istr, buf));
            ORC_PEEKC;
            if (endiow()) {
              if (parse_weight(buf.c_str(), weight)) {
                inL = outL = EPS;
              } else {
                inL = in.index_of(buf);
                outL = out.index_of(buf);
                weight = W::one();
              }
            } else {
              inL = in.index_of(buf);
              ORC_REQ(getString(istr, buf2));
              ORC_PEEKC;
              if (endiow()) {
                if (parse_weight(buf2.c_str(), weight)) {
                  outL = out.index_of(buf);
                } else {
                  outL = out.index_of(buf2);
                  weight = W::one();
                }
              } else {
                outL = out.index_of(buf2);
                std::string wtok;
                ORC_REQ(getString(istr, wtok));
                ORC_REQ(parse_weight(wtok.c_str(), weight));
                ORC_PEEKC;
                ORC_REQ(endiow());
              }
            }
          }
          Arc a{inL, outL, dst, weight, NO_GROUP};
          ORC_GETC;
          if (c == '!') {
            ORC_PEEKC;
            if (isdigit((unsigned char)c)) {
              unsigned g;
              ORC_REQ(istr >> g);
              a.group = g;
            } else
              a.group = LOCKED_GROUP;
          } else
            istr.unget();
          states[src].push_back(a);
          if (!iowparen) break;
          ORC_REQ(istr >> c && c == ')');
          ORC_PEEKC;
          if (c == ')') break;
        }
        if (!destparen) break;
        ORC_REQ(istr >> c && c == ')');
      }
      ORC_REQ(istr >> c && c == ')');
    }
#undef ORC_REQ
#undef ORC_GETC
#undef ORC_PEEKC
    if (!named) {
      if (!(final_state < states.size())) return false;
      is_valid = true;
      return true;
    }
    auto it = stateIdx.find(finalName);
    if (it == stateIdx.end()) return false;
    final_state = it->second;
    is_valid = true;
    return true;
  }
  bool read_file(std::string const& path, bool alwaysNamed = true) {
    std::ifstream f(path);
    if (!f) return false;
    return read(f, alwaysNamed);
  }

  // ---- writer: wfstio.cc:594-625 (brief/full, state-per-line/arc-per-line) ----
  void write(std::ostream& os, bool full = false, bool onearc = false, bool include_zero = false,
             WeightFormat const& wf = WeightFormat()) const {
    bool brief = !full;
    if (!valid()) return;
    os << stateName(final_state);
    for (unsigned i = 0; i < numStates(); ++i) {
      if (!onearc) os << "\n(" << stateName(i);
      for (Arc const& a : states[i]) {
        if (include_zero || a.weight.isPositive()) {
          if (onearc) os << "\n(" << stateName(i);
          os << " (" << stateName(a.dest);
          if (!brief || a.in || a.out) {
            std::string const &il = alph[0]->names[a.in], &ol = alph[1]->names[a.out];
            os << " " << il;
            if (!brief || il != ol) os << " " << ol;
          }
          if (!brief || ~a.group || a.weight != W::one()) os << " " << fmt_weight(a.weight, wf);
          if (~a.group) {
            os << '!';
            if (a.group > 0) os << a.group;
          }
          os << ")";
          if (onearc) os << ")";
        }
      }
      if (!onearc) os << ")";
    }
    os << "\n";
  }

  // ---- reduce: carmel/src/fst.cc:468-545 (+ state.h:280-314) ----
  void reduce() {
    unsigned n = numStates();
    if (!valid()) {
      states.clear();
      return;
    }
    std::vector<char> fwd(n, 0), bwd(n, 0);
    std::vector<std::vector<unsigned>> rev(n);
    for (unsigned s = 0; s < n; ++s)
      for (Arc const& a : states[s]) rev[a.dest].push_back(s);
    std::vector<unsigned> stack;
    stack.push_back(0);
    fwd[0] = 1;
    while (!stack.empty()) {
      unsigned s = stack.back();
      stack.pop_back();
      for (Arc const& a : states[s])
        if (!fwd[a.dest]) {
          fwd[a.dest] = 1;
          stack.push_back(a.dest);
        }
    }
    stack.push_back(final_state);
    bwd[final_state] = 1;
    while (!stack.empty()) {
      unsigned s = stack.back();
      stack.pop_back();
      for (unsigned p : rev[s])
        if (!bwd[p]) {
          bwd[p] = 1;
          stack.push_back(p);
        }
    }
    std::vector<unsigned> oldToNew(n);
    unsigned k = 0;
    bool removed = false;
    for (unsigned i = 0; i < n; ++i) {
      if (fwd[i] && bwd[i])
        oldToNew[i] = k++;
      else {
        oldToNew[i] = ~0u;
        removed = true;
      }
    }
    if (removed) {  // fst.cc:530-545 removeMarkedStates: relative order kept
      std::vector<Arcs> ns(k);
      std::vector<std::string> nn;
      for (unsigned i = 0; i < n; ++i)
        if (~oldToNew[i]) {
          ns[oldToNew[i]].swap(states[i]);
          if (named && i < stateNames.size()) nn.push_back(stateNames[i]);
        }
      states.swap(ns);
      if (named) {
        stateNames.swap(nn);
        stateIdx.clear();
        for (unsigned i = 0; i < stateNames.size(); ++i) stateIdx.emplace(stateNames[i], i);
      }
      for (auto& st : states) {
        Arcs keep;
        for (Arc& a : st)
          if (~oldToNew[a.dest]) {
            a.dest = oldToNew[a.dest];
            keep.push_back(a);
          }
        st.swap(keep);
      }
      if (!~oldToNew[final_state]) {
        is_valid = false;
        states.clear();
        return;
      }
      final_state = oldToNew[final_state];
    }
    for (unsigned i = 0; i < numStates(); ++i) {  // remove_epsilons_to(i), state.h:280-289
      Arcs keep;
      for (Arc const& a : states[i])
        if (!(a.in == 0 && a.out == 0 && a.dest == i)) keep.push_back(a);
      states[i].swap(keep);
    }
  }

  // ---- normalize: carmel/src/fst.cc:86-244; group enumeration fst.h:1362-1446 ----
  // CONDITIONAL groups iterate the per-state by-input hash index whose lists hold arcs in
  // reverse arc order (state.h:195-199 push_front); the hash-bucket order of groups is
  // implementation defined and only affects rounding of tie-group sums, so groups are
  // enumerated here by first occurrence.
  void norm_groups(NormGroupBy g, std::vector<std::vector<Arc*>>& groups) {
    groups.clear();
    for (auto& st : states) {
      if (g == JOINT) {
        groups.emplace_back();
        for (Arc& a : st) groups.back().push_back(&a);
      } else {
        if (st.empty()) continue;
        std::unordered_map<unsigned, size_t> byin;
        size_t base = groups.size();
        for (auto it = st.rbegin(); it != st.rend(); ++it) {
          auto f = byin.find(it->in);
          size_t gi;
          if (f == byin.end()) {
            gi = groups.size();
            byin.emplace(it->in, gi);
            groups.emplace_back();
          } else
            gi = f->second;
          groups[gi].push_back(&*it);
        }
        (void)base;
      }
    }
  }
  void normalize(NormalizeMethod const& method, bool uniform_zero_normgroups = false) {
    if (method.group == NONE) return;
    std::vector<std::vector<Arc*>> groups;
    norm_groups(method.group, groups);
    W addc = method.add_count;
    std::unordered_map<unsigned, W> groupArcTotal, groupStateTotal, groupMaxLockedSum;
    for (auto& g : groups) {  // pass 1, fst.cc:115-153
      W sum, locked_sum;
      for (Arc* a : g) {
        a->weight += addc;
        if (a->isLocked())
          locked_sum += a->weight;
        else
          sum += a->weight;
      }
      for (Arc* a : g)
        if (a->isTied()) {
          groupArcTotal[a->group] += a->weight;
          groupStateTotal[a->group] += sum;
          W& m = groupMaxLockedSum[a->group];
          if (locked_sum > m) m = locked_sum;
        }
    }
    for (auto& g : groups) {  // pass 2, fst.cc:160-229
      W normal_sum, reserved;
      for (Arc* a : g) {
        if (a->isTied()) {
          W groupNorm = groupStateTotal[a->group];
          W gmax = groupMaxLockedSum[a->group];
          W one = W::one();
          if (gmax > one) {
            a->weight.setZero();
          } else {
            if (!gmax.isZero()) groupNorm /= (one - gmax);
            W groupTotal = groupArcTotal[a->group];
            if (!groupTotal.isZero()) {
              a->weight = groupTotal / groupNorm;
              reserved += a->weight;
            } else
              a->weight.setZero();
          }
        } else if (a->isLocked()) {
          reserved += a->weight;
        } else {
          normal_sum += a->weight;
        }
      }
      W fraction_remain = W::one();
      fraction_remain -= reserved;
      bool something_left = !fraction_remain.isZero();
      if (something_left && (uniform_zero_normgroups || !normal_sum.isZero())) {
        for (Arc* a : g)
          if (a->isNormal()) a->weight = fraction_remain * a->weight / normal_sum;
      } else
        for (Arc* a : g)
          if (a->isNormal()) a->weight.setZero();
    }
  }

  // fst.h:986-988 zero_arcs via state.h:84-103 modify_parameter_once (locked arcs untouched)
  void zero_arcs() {
    for (auto& st : states)
      for (Arc& a : st)
        if (!a.isLocked()) a.weight.setZero();
  }
  template <class V>
  void visit_arcs(V&& v) {  // fst.h:1330-1334: state 0..n, list order
    for (unsigned s = 0; s < numStates(); ++s)
      for (Arc& a : states[s]) v(s, a);
  }
};

// ------------------------------------------------------------------------------------------
// training corpus  (carmel/src/train.h:60-189, train.cc:985-1025, wfstio.cc:631-651)
// ------------------------------------------------------------------------------------------
struct Example {
  std::vector<unsigned> in, out;
  double weight = 1;
};
struct Corpus {
  std::list<Example> examples;
  unsigned n_pairs = 0;
  double totalEmpiricalWeight = 0, n_input = 0, n_output = 0;
  void count() {  // train.h:147-163
    n_pairs = 0;
    totalEmpiricalWeight = n_input = n_output = 0;
    for (auto const& e : examples) {
      n_input += e.in.size();
      n_output += e.out.size();
      totalEmpiricalWeight += e.weight;
      ++n_pairs;
    }
  }
  static void symbol_list(std::vector<unsigned>& ret, std::string const& line, Alphabet& a) {
    std::istringstream is(line);
    std::string sym;
    while (is) {
      if (!WFST::getString(is, sym)) break;
      ret.push_back(a.index_of(sym));
    }
  }
  void read(std::istream& in, WFST& x) {
    std::string buf;
    for (;;) {
      double weight = 1;
      if (!std::getline(in, buf)) break;
      char s = buf.empty() ? 0 : buf[0];
      if (isdigit((unsigned char)s) || s == '-' || s == '.' || s == 'e') {
        std::istringstream w(buf);
        if (!(w >> weight)) continue;
        if (!std::getline(in, buf)) break;
      }
      Example e;
      e.weight = weight;
      symbol_list(e.in, buf, *x.alph[0]);
      if (!std::getline(in, buf)) break;
      symbol_list(e.out, buf, *x.alph[1]);
      examples.push_back(std::move(e));
    }
    count();
  }
};

// ------------------------------------------------------------------------------------------
// cascade_parameters  (carmel/src/cascade.h:22-671)
// ------------------------------------------------------------------------------------------
struct Cascade {
  bool trivial = true;
  std::vector<WFST*> cascade;
  typedef std::vector<Arc*> chain_t;  // list order = slist order (cons prepends)
  std::vector<chain_t> chains;
  std::vector<W> chain_weights;
  std::unordered_map<Arc*, unsigned> epsilon_chains;
  unsigned nil_chain = 0;
  WFST* pcomposed = nullptr;
  bool is_chain[2] = {false, false};

  explicit Cascade(bool remember = false) {  // cascade.h:366-383
    trivial = !remember;
    if (trivial) return;
    nil_chain = 0;
    chains.emplace_back();
  }
  void set_trivial() {
    chain_weights.clear();
    epsilon_chains.clear();
    trivial = true;
  }
  void add(WFST* w) {
    if (!trivial) cascade.push_back(w);
  }
  void set_composed(WFST* c) {  // cascade.h:205-208
    pcomposed = c;
    if (trivial) cascade.assign(1, c);
  }
  void prepare_compose(bool first_chain, bool second_chain) {
    is_chain[0] = first_chain;
    is_chain[1] = second_chain;
  }
  static bool is_locked_1(Arc* e) { return e->isLocked() && e->weight.isOne(); }
  chain_t cons(Arc* a, chain_t cdr) {  // cascade.h:511-514
    if (is_locked_1(a)) return cdr;
    cdr.insert(cdr.begin(), a);
    return cdr;
  }
  chain_t cons_chain(Arc* a, Arc* b) {  // cascade.h:544-559
    if (is_chain[0]) {
      chain_t ca = chains[a->group];
      if (is_chain[1]) {
        chain_t r = chains[b->group];
        for (Arc* x : ca) r = cons(x, r);
        return r;
      }
      return cons(b, ca);
    } else {
      if (is_chain[1]) return cons(a, chains[b->group]);
      return cons(a, cons(b, chain_t()));
    }
  }
  unsigned original_id(Arc* e) { return is_locked_1(e) ? nil_chain : e->group; }
  unsigned record_eps(Arc* e, bool chain) {  // cascade.h:573-586
    if (trivial) return e->group;
    if (chain) return original_id(e);
    auto ins = epsilon_chains.emplace(e, (unsigned)chains.size());
    if (ins.second) {
      chain_t v = cons(e, chain_t());
      if (v.empty()) return ins.first->second = nil_chain;
      chains.push_back(v);
    }
    return ins.first->second;
  }
  unsigned record1(Arc* e) { return record_eps(e, is_chain[0]); }
  unsigned record2(Arc* e) { return record_eps(e, is_chain[1]); }
  unsigned record(Arc* a, Arc* b) {  // cascade.h:588-599
    if (trivial) return NO_GROUP;
    chain_t v = cons_chain(a, b);
    if (v.empty()) return nil_chain;
    chains.push_back(v);
    return (unsigned)chains.size() - 1;
  }
  unsigned locked_1_groupid() { return trivial ? LOCKED_GROUP : nil_chain; }
  void done_composing(WFST* composed) {
    set_composed(composed);
    if (trivial) return;
    epsilon_chains.clear();
  }
  void normalize(std::vector<NormalizeMethod> const& m) {  // cascade.h:385-388
    for (unsigned i = 0; i < cascade.size(); ++i) cascade[i]->normalize(m[i]);
  }
  void calculate_chain_weights() {  // cascade.h:426-433
    chain_weights.assign(chains.size(), W::one());
    for (unsigned i = 0; i < chains.size(); ++i)
      for (Arc* p : chains[i]) chain_weights[i] *= p->weight;
  }
  void update() {  // cascade.h:466-479
    if (trivial) return;
    calculate_chain_weights();
    for (auto& st : pcomposed->states)
      for (Arc& a : st) a.weight = chain_weights[a.group];
  }
  void distribute_counts() {  // cascade.h:286-325
    if (trivial) return;
    for (WFST* w : cascade) w->zero_arcs();
    for (auto& st : pcomposed->states)
      for (Arc& a : st)
        for (Arc* p : chains[a.group])
          if (!p->isLocked()) p->weight += a.weight;
  }
  std::vector<std::vector<W>> none_saves;
  void save_none(std::vector<NormalizeMethod> const& m) {  // cascade.h:339-343
    none_saves.assign(m.size(), {});
    for (unsigned i = 0; i < std::min(m.size(), cascade.size()); ++i)
      if (m[i].group == NONE) cascade[i]->visit_arcs([&](unsigned, Arc& a) { none_saves[i].push_back(a.weight); });
  }
  void load_none(std::vector<NormalizeMethod> const& m) {  // cascade.h:345-350
    for (unsigned i = 0; i < std::min(m.size(), cascade.size()); ++i)
      if (m[i].group == NONE) {
        size_t k = 0;
        cascade[i]->visit_arcs([&](unsigned, Arc& a) { a.weight = none_saves[i][k++]; });
        none_saves[i].clear();
      }
  }
  void use_counts(std::vector<NormalizeMethod> const& m) {  // cascade.h:353-356
    distribute_counts();
    normalize(m);
  }
  void use_counts_final(std::vector<NormalizeMethod> const& m) {  // cascade.h:358-364
    if (trivial) return;
    save_none(m);
    use_counts(m);
    load_none(m);
    update();
  }
};

// ------------------------------------------------------------------------------------------
// composition with the 3-state epsilon filter  (carmel/src/compose.cc:163-531, non "-a" path)
// Composed-state numbering = discovery order with a LIFO work list (compose.cc:193,215,326-328;
// list.h:119-128), arcs pushed to the FRONT of the source state's list (compose.cc:139).
// The by-label index lists hold arcs in reverse arc order (state.h:195-199); the per-state
// indexing threshold is WFST::indexThreshold (carmel.cc -T, default 32, carmel.cc:895,1078).
// ------------------------------------------------------------------------------------------
struct ComposeResult {
  std::unique_ptr<WFST> fst;
};
inline std::unique_ptr<WFST> compose(Cascade& cascade, WFST& a, WFST& b, unsigned indexThreshold = 32) {
  std::unique_ptr<WFST> rp(new WFST());
  WFST& r = *rp;
  r.alph[0] = a.alph[0];
  r.alph[1] = b.alph[1];
  r.named = false;
  if (!(a.valid() && b.valid())) return rp;
  Alphabet &aout = *a.alph[1], &bin = *b.alph[0];
  std::vector<unsigned> map(aout.names.size()), revMap(bin.names.size());
  // alphabet computeMap: matching symbol by string, else no match (use ~0u-1 so nothing matches)
  for (unsigned i = 0; i < aout.names.size(); ++i) {
    int j = bin.find(aout.names[i]);
    map[i] = j < 0 ? 0xFFFFFFFEu : (unsigned)j;
  }
  for (unsigned i = 0; i < bin.names.size(); ++i) {
    int j = aout.find(bin.names[i]);
    revMap[i] = j < 0 ? 0xFFFFFFFEu : (unsigned)j;
  }
  struct Trio {
    unsigned qa, qb;
    char f;
  };
  auto key = [&](Trio const& t) { return ((uint64_t)t.f * a.numStates() + t.qa) * (uint64_t)b.numStates() + t.qb; };
  std::unordered_map<uint64_t, unsigned> stateMap;
  std::vector<std::pair<unsigned, Trio>> queue;  // LIFO
  Trio t0{0, 0, 0};
  stateMap[key(t0)] = 0;
  r.states.emplace_back();
  queue.push_back({0, t0});
  // per-state label index, built lazily: label -> arcs in REVERSE arc order
  typedef std::unordered_map<unsigned, std::vector<Arc*>> Index;
  std::vector<std::unique_ptr<Index>> aIdx(a.numStates()), bIdx(b.numStates());
  auto indexBy = [&](WFST& w, std::vector<std::unique_ptr<Index>>& idx, unsigned s, bool byOut) -> Index& {
    if (!idx[s]) {
      idx[s].reset(new Index());
      auto& st = w.states[s];
      for (auto it = st.rbegin(); it != st.rend(); ++it) (*idx[s])[byOut ? it->out : it->in];
      for (auto it = st.rbegin(); it != st.rend(); ++it) (*idx[s])[byOut ? it->out : it->in].push_back(&*it);
    }
    return *idx[s];
  };
  unsigned sourceState = 0;
  auto composeArc = [&](unsigned in, unsigned out, Trio const& dest, W weight, unsigned g) {
    auto ins = stateMap.emplace(key(dest), r.numStates());
    unsigned num;
    if (ins.second) {
      num = r.numStates();
      queue.push_back({num, dest});
      r.states.emplace_back();
    } else
      num = ins.first->second;
    r.states[sourceState].push_front(Arc{in, out, num, weight, g});
  };
  while (!queue.empty()) {
    sourceState = queue.back().first;
    Trio src = queue.back().second;
    queue.pop_back();
    auto &qa = a.states[src.qa], &qb = b.states[src.qb];
    bool larger_is_a = qa.size() > qb.size();
    size_t larger_size = larger_is_a ? qa.size() : qb.size();
    Trio d;
    if (larger_size > indexThreshold) {
      if (!larger_is_a) {  // qb larger: compose.cc:338-392
        Index& bi = indexBy(b, bIdx, src.qb, false);
        for (Arc& l : qa) {
          unsigned in = l.in;
          d.qa = l.dest;
          if (l.out == EPS) {
            if (src.f != 2) {
              d.f = 1;
              d.qb = src.qb;
              composeArc(in, EPS, d, l.weight, cascade.record1(&l));
            }
            if (src.f == 0) {
              auto m = bi.find(EPS);
              if (m != bi.end()) {
                d.f = 0;
                for (Arc* rr : m->second) {
                  d.qb = rr->dest;
                  composeArc(in, rr->out, d, l.weight * rr->weight, cascade.record(&l, rr));
                }
              }
            }
          } else {
            auto m = bi.find(map[l.out]);
            if (m != bi.end()) {
              d.f = 0;
              for (Arc* rr : m->second) {
                d.qb = rr->dest;
                composeArc(in, rr->out, d, l.weight * rr->weight, cascade.record(&l, rr));
              }
            }
          }
        }
        if (src.f != 1) {
          auto m = bi.find(EPS);
          if (m != bi.end()) {
            d.qa = src.qa;
            d.f = 2;
            for (Arc* rr : m->second) {
              d.qb = rr->dest;
              composeArc(EPS, rr->out, d, rr->weight, cascade.record2(rr));
            }
          }
        }
      } else {  // qa larger: compose.cc:393-445
        Index& ai = indexBy(a, aIdx, src.qa, true);
        for (Arc& rr : qb) {
          unsigned out = rr.out;
          d.qb = rr.dest;
          if (rr.in == EPS) {
            if (src.f != 1) {
              d.f = 2;
              d.qa = src.qa;
              composeArc(EPS, out, d, rr.weight, cascade.record2(&rr));
            }
            if (src.f == 0) {
              auto m = ai.find(EPS);
              if (m != ai.end()) {
                d.f = 0;
                for (Arc* l : m->second) {
                  d.qa = l->dest;
                  composeArc(l->in, out, d, l->weight * rr.weight, cascade.record(l, &rr));
                }
              }
            }
          } else {
            d.f = 0;
            auto m = ai.find(revMap[rr.in]);
            if (m != ai.end())
              for (Arc* l : m->second) {
                d.qa = l->dest;
                composeArc(l->in, out, d, l->weight * rr.weight, cascade.record(l, &rr));
              }
          }
        }
        if (src.f != 2) {
          auto m = ai.find(EPS);
          if (m != ai.end()) {
            d.qb = src.qb;
            d.f = 1;
            for (Arc* l : m->second) {
              d.qa = l->dest;
              composeArc(l->in, EPS, d, l->weight, cascade.record1(l));
            }
          }
        }
      }
    } else {  // compose.cc:446-497
      for (Arc& l : qa) {
        unsigned in = l.in;
        d.qa = l.dest;
        if (l.out == EPS) {
          if (src.f != 2) {
            d.f = 1;
            d.qb = src.qb;
            composeArc(in, EPS, d, l.weight, cascade.record1(&l));
          }
          if (src.f == 0) {
            for (Arc& rr : qb)
              if (rr.in == EPS) {
                d.qb = rr.dest;
                d.f = 0;
                composeArc(in, rr.out, d, l.weight * rr.weight, cascade.record(&l, &rr));
              }
          }
        } else {
          d.f = 0;
          for (Arc& rr : qb)
            if (map[l.out] == rr.in) {
              d.qb = rr.dest;
              composeArc(in, rr.out, d, l.weight * rr.weight, cascade.record(&l, &rr));
            }
        }
      }
      if (src.f != 1) {
        d.qa = src.qa;
        d.f = 2;
        for (Arc& rr : qb)
          if (rr.in == EPS) {
            d.qb = rr.dest;
            composeArc(EPS, rr.out, d, rr.weight, cascade.record2(&rr));
          }
      }
    }
  }
  // finals: compose.cc:503-528
  unsigned nFinal = 0;
  unsigned pFinal[3];
  bool has[3] = {false, false, false};
  for (int i = 0; i < 3; ++i) {
    Trio t{a.final_state, b.final_state, (char)i};
    auto it = stateMap.find(key(t));
    if (it != stateMap.end()) {
      has[i] = true;
      pFinal[i] = it->second;
      ++nFinal;
      r.final_state = it->second;
    }
  }
  if (nFinal == 0) return rp;
  if (nFinal > 1) {
    r.final_state = r.numStates();
    r.states.emplace_back();
    for (int i = 0; i < 3; ++i)
      if (has[i]) r.states[pFinal[i]].push_front(Arc{EPS, EPS, r.final_state, W::one(), cascade.locked_1_groupid()});
  }
  r.is_valid = true;
  return rp;
}

// ------------------------------------------------------------------------------------------
// derivations: per-example trellis (carmel/src/derivations.h:45-66,79-155,479-513,572-704)
// graph arcs: graehl/shared/graph.h:37-121 ; ordering :241-288 ; reverse graph.cc:41-57
// ------------------------------------------------------------------------------------------
struct GraphArc {
  unsigned src, dest;
  double weight;  // real weight at build time (derivations.h:698), only used by gibbs init
  unsigned id;  // arcs_table id
};
typedef std::list<GraphArc> GArcs;  // List<GraphArc>, push_front on add (graph.h:91-94)

struct ArcTableEntry {  // train.h:28-40 arc_counts
  Arc* arc;
  unsigned src;
  W scratch, em_weight, best_weight, counts, prior_counts;
};
struct ArcsTable {  // derivations.h:79-101: id = visit order
  std::vector<ArcTableEntry> t;
  unsigned n_states;
  ArcsTable(WFST& x, bool per_arc_prior, W global_prior) {
    n_states = x.numStates();
    x.visit_arcs([&](unsigned s, Arc& a) {
      ArcTableEntry e;
      e.arc = &a;
      e.src = s;
      e.prior_counts = per_arc_prior ? global_prior + a.weight : global_prior;
      t.push_back(e);
    });
  }
  size_t size() const { return t.size(); }
};
struct IOIndex {  // derivations.h:142-155
  std::vector<std::unordered_map<uint64_t, std::vector<unsigned>>> st;
  explicit IOIndex(WFST& x) : st(x.numStates()) {
    unsigned i = 0;
    x.visit_arcs([&](unsigned s, Arc& a) { st[s][((uint64_t)a.in << 32) | a.out].push_back(i++); });
  }
};

struct DerivStateKey {
  uint32_t i, s, o;
  bool operator==(DerivStateKey const& r) const { return i == r.i && s == r.s && o == r.o; }
};
struct DerivStateHash {
  size_t operator()(DerivStateKey const& k) const {
    uint64_t h = k.i * 0x9E3779B97F4A7C15ull;
    h ^= (k.s + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= (h >> 29);
    h += k.o * 0x165667B19E3779F9ull;
    return (size_t)(h ^ (h >> 32));
  }
};

struct Derivations {
  std::vector<unsigned> in, out;
  std::vector<GArcs> g;
  unsigned fin = 0;
  bool no_goal = true;
  double weight = 1;
  std::unordered_map<DerivStateKey, unsigned, DerivStateHash> id_of_state;
  std::vector<char> remove;
  DerivStateKey goal;
  size_t pre_arcs = 0;  // global_stats.pre.arcs contribution

  bool empty() const { return no_goal; }
  size_t n_states() const { return g.size(); }
  size_t n_arcs() const {
    size_t n = 0;
    for (auto const& s : g) n += s.size();
    return n;
  }

  // derivations.h:640-675
  unsigned derive(IOIndex const& io, ArcsTable const& atab, DerivStateKey const& d) {
    unsigned src = (unsigned)g.size();
    auto ins = id_of_state.emplace(d, src);
    if (!ins.second) return ins.first->second;
    g.emplace_back();
    remove.push_back(0);
    auto const& fs = io.st[d.s];
    bool dead = !(d == goal);
    if (add_arcs(io, atab, EPS, EPS, d.i, d.o, fs, src)) dead = false;
    bool useO = d.o < out.size(), useI = d.i < in.size();
    unsigned o1 = d.o + 1, i1 = d.i + 1;
    if (useO)
      if (add_arcs(io, atab, EPS, out[d.o], d.i, o1, fs, src)) dead = false;
    if (useI) {
      unsigned si = in[d.i];
      if (add_arcs(io, atab, si, EPS, i1, d.o, fs, src)) dead = false;
      if (useO)
        if (add_arcs(io, atab, si, out[d.o], i1, o1, fs, src)) dead = false;
    }
    remove[src] = dead;
    return src;
  }
  // derivations.h:678-704
  bool add_arcs(IOIndex const& io, ArcsTable const& atab, unsigned s_in, unsigned s_out, unsigned i_in,
                unsigned i_out, std::unordered_map<uint64_t, std::vector<unsigned>> const& fs, unsigned source) {
    bool reachgoal = false;
    auto m = fs.find(((uint64_t)s_in << 32) | s_out);
    if (m != fs.end())
      for (unsigned id : m->second) {
        ++pre_arcs;
        Arc* a = atab.t[id].arc;
        DerivStateKey ds{i_in, a->dest, i_out};
        unsigned dst = derive(io, atab, ds);
        if (!remove[dst]) {
          g[source].push_front(GraphArc{source, dst, a->weight.getReal(), id});
          reachgoal = true;
        }
      }
    return reachgoal;
  }
  // derivations.h:572-629 + array.hpp:73-92 + graph.h:318-345
  void prune() {
    if (empty()) return;
    unsigned n = (unsigned)g.size();
    std::vector<unsigned> ttable(n);
    unsigned k = 0;
    for (unsigned i = 0; i < n; ++i) ttable[i] = remove[i] ? ~0u : k++;
    remove.clear();
    fin = ttable[fin];
    std::vector<GArcs> ng(k);
    for (unsigned i = 0; i < n; ++i) {
      if (!~ttable[i]) continue;
      GArcs& arcs = g[i];
      for (auto it = arcs.begin(); it != arcs.end();) {
        if (~ttable[it->dest]) {
          it->src = ttable[it->src];
          it->dest = ttable[it->dest];
          ++it;
        } else
          it = arcs.erase(it);
      }
      ng[ttable[i]].swap(arcs);
    }
    g.swap(ng);
  }
  // derivations.h:479-513
  bool compute(WFST& x, IOIndex const& io, ArcsTable const& atab, bool prune_ = true) {
    remove.clear();
    id_of_state.clear();
    g.clear();
    goal = DerivStateKey{(uint32_t)in.size(), x.final_state, (uint32_t)out.size()};
    derive(io, atab, DerivStateKey{0, 0, 0});
    auto pf = id_of_state.find(goal);
    no_goal = (pf == id_of_state.end());
    if (!no_goal) fin = pf->second;
    // note: the reference records the goal even if it was marked dead?  goal is never dead
    // (derivations.h:655 dead = (d != goal)).
    if (prune_) prune();
    id_of_state.clear();
    if (no_goal) {
      g.clear();
      return false;
    }
    return true;
  }

  // graph.h:241-288 reverse_topo_order::order_from (recursive DFS post-order), made iterative-safe
  // by explicit recursion here (depth is bounded by the trellis depth like the reference).
  void make_order(std::vector<unsigned>& reverse_order, unsigned* n_back_edges = nullptr) const {
    unsigned n = (unsigned)g.size();
    std::vector<char> done(n, 0), begun(n, 0);
    reverse_order.clear();
    reverse_order.reserve(n);
    unsigned nback = 0;
    struct Frame {
      unsigned s;
      GArcs::const_iterator it;
    };
    std::vector<Frame> stack;
    auto enter = [&](unsigned s) -> bool {
      if (done[s]) return false;
      if (begun[s]) {
        ++nback;
        return false;
      }
      begun[s] = 1;
      stack.push_back(Frame{s, g[s].begin()});
      return true;
    };
    enter(0);
    while (!stack.empty()) {
      Frame& f = stack.back();
      if (f.it == g[f.s].end()) {
        done[f.s] = 1;
        reverse_order.push_back(f.s);
        stack.pop_back();
      } else {
        unsigned d = f.it->dest;
        ++f.it;
        enter(d);
      }
    }
    if (n_back_edges) *n_back_edges = nback;
  }
  // graph.cc:41-57 add_reversed_arcs: push_front onto rev[dest]
  void make_reverse(std::vector<GArcs>& r) const {
    r.assign(g.size(), GArcs());
    for (unsigned i = 0; i < g.size(); ++i)
      for (GraphArc const& a : g[i]) r[a.dest].push_front(GraphArc{a.dest, a.src, a.weight, a.id});
  }

  // derivations.h:400-417 compute_fb with weight functor wt(id)
  template <class WF>
  W compute_fb(std::vector<W>& f, std::vector<W>& b, WF const& wt) const {
    unsigned nst = (unsigned)g.size();
    f.assign(nst, W());
    b.assign(nst, W());
    f[0] = W::one();
    std::vector<unsigned> reverse_order;
    make_order(reverse_order);
    for (auto t = reverse_order.rbegin(); t != reverse_order.rend(); ++t) {  // graph.h:391-402
      unsigned src = *t;
      for (GraphArc const& a : g[src]) f[a.dest] += f[src] * wt(a);
    }
    W prob = f[fin];
    std::vector<GArcs> r;
    make_reverse(r);
    b[fin] = W::one();
    for (auto t = reverse_order.begin(); t != reverse_order.end(); ++t) {
      unsigned src = *t;
      for (GraphArc const& a : r[src]) b[a.dest] += b[src] * wt(a);
    }
    return prob;
  }
  // derivations.h:432-449
  W collect_counts(ArcsTable& t) const {
    std::vector<W> f, b;
    auto wt = [&](GraphArc const& a) { return t.t[a.id].arc->weight; };
    W prob = compute_fb(f, b, wt);
    unsigned nst = (unsigned)g.size();
    for (unsigned s = 0; s < nst; ++s)
      for (GraphArc const& a : g[s]) {
        ArcTableEntry& ac = t.t[a.id];
        W arc_contrib = ac.arc->weight * f[a.src] * b[a.dest];
        ac.counts += arc_contrib * W(weight) / prob;
      }
    return prob;
  }
};

// ------------------------------------------------------------------------------------------
// EM training loop  (carmel/src/train.cc:119-221,503-678,763-773,893-923; cached_derivs.h:60-101)
// ------------------------------------------------------------------------------------------
struct TrainOpts {
  unsigned max_iter = 500;  // fst.h:1089
  W converge_arc_delta = W(1e-4);  // carmel.cc:896
  W converge_perplexity_ratio = W(.999);  // carmel.cc:897
  W smoothFloor;  // -f
  bool weight_is_prior_count = false;  // -U
  double learning_rate_growth_factor = 1.;  // -o
  bool cache_derivations = false;  // -: (cache_forward_backward); default re-derives (fst.h:1072)
  bool prune = true;
  bool quiet = false;
};
struct IterRecord {
  unsigned iter;
  double ln_prob;  // unweighted corpus log prob (ln)
  double ln_weighted_prob;
  double max_change;
};
struct Trainer {
  WFST& x;
  Cascade& cascade;
  Corpus& corpus;
  std::vector<NormalizeMethod> methods;
  TrainOpts opt;
  ArcsTable arcs;
  std::ostream& log;
  bool first = true;
  std::vector<Derivations> cache;
  bool cached = false;
  std::vector<IterRecord> history;
  size_t total_trellis_arcs = 0, total_trellis_states = 0;  // of the last E-step

  Trainer(WFST& x, Cascade& c, Corpus& corpus, std::vector<NormalizeMethod> const& m, TrainOpts const& o,
          std::ostream& log)
      : x(x), cascade(c), corpus(corpus), methods(m), opt(o), arcs(initial_normalize(x, c, m), o.weight_is_prior_count, o.smoothFloor), log(log) {}
  // train.cc:508-513: set_composed + cascade.normalize(methods) happen BEFORE the arcs table is built
  // (so -U per-arc priors see the normalised weights)
  static WFST& initial_normalize(WFST& x, Cascade& c, std::vector<NormalizeMethod> const& m) {
    c.set_composed(&x);
    c.normalize(m);
    return x;
  }

  // cached_derivs.h:60-101 foreach_deriv + train.cc:326-332 functor
  W estimate(W& unweighted) {
    for (auto& a : arcs.t) a.counts.setZero();  // train.cc:764
    unweighted = W::one();
    W weighted = W::one();
    total_trellis_arcs = total_trellis_states = 0;
    if (opt.cache_derivations) {
      if (!cached) {
        IOIndex io(x);
        for (auto i = corpus.examples.begin(); i != corpus.examples.end();) {
          Derivations d;
          d.in = i->in;
          d.out = i->out;
          d.weight = i->weight;
          if (d.compute(x, io, arcs, opt.prune)) {
            d.in.clear();
            d.out.clear();
            cache.push_back(std::move(d));
            ++i;
          } else {
            if (!opt.quiet) log << "No derivations in transducer for input/output\n";
            i = corpus.examples.erase(i);
          }
        }
        corpus.count();
        cached = true;
      }
      for (auto& d : cache) {
        W prob = d.collect_counts(arcs);
        unweighted *= prob;
        weighted *= prob.pow(d.weight);
        total_trellis_arcs += d.n_arcs();
        total_trellis_states += d.n_states();
      }
    } else {
      IOIndex io(x);
      unsigned n = 0;
      for (auto i = corpus.examples.begin(); i != corpus.examples.end();) {
        ++n;
        Derivations d;
        d.in = i->in;
        d.out = i->out;
        d.weight = i->weight;
        if (d.compute(x, io, arcs, opt.prune)) {
          W prob = d.collect_counts(arcs);
          unweighted *= prob;
          weighted *= prob.pow(d.weight);
          total_trellis_arcs += d.n_arcs();
          total_trellis_states += d.n_states();
        } else if (first) {
          if (!opt.quiet) log << "No derivations in transducer for input/output #" << n << "\n";
          if (opt.prune) {
            i = corpus.examples.erase(i);
            continue;
          }
        }
        ++i;
      }
      if (first) corpus.count();
    }
    first = false;
    if (corpus.examples.empty()) throw std::runtime_error("No training example had a derivation - aborting training.");
    return weighted;
  }

  // train.cc:893-923
  W maximize(double delta_scale) {
    cascade.save_none(methods);
    for (auto& a : arcs.t)  // prep_new_weights(1.0), train.cc:134-153
      if (!a.arc->isLocked()) {
        a.scratch = a.arc->weight;
        a.arc->weight = a.counts + a.prior_counts * W(1.0);
      }
    if (cascade.trivial)
      x.normalize(methods[0]);  // cascade.use_counts: distribute (no-op) + normalize cascade[0]==x
    else
      cascade.use_counts(methods);
    cascade.load_none(methods);
    if (cascade.trivial) {
      for (auto& a : arcs.t) {  // overrelax, train.cc:157-171
        a.em_weight = a.arc->weight;
        if (delta_scale > 1.)
          if (!a.arc->isLocked())
            if (a.scratch.isPositive()) a.arc->weight = a.scratch * ((a.em_weight / a.scratch).pow(delta_scale));
      }
      if (delta_scale > 1.) x.normalize(methods[0]);
      W maxChange;
      for (auto& a : arcs.t)
        if (!a.arc->isLocked()) {
          W change = absdiff(a.arc->weight, a.scratch);
          if (change > maxChange) maxChange = change;
        }
      return maxChange;
    }
    return W(10);
  }

  void print_ppx(W corpus_p) {  // weight.h:311-329 print_ppx_symbol
    log << "probability=" << fmt_base2(corpus_p);
    double n_symbol = std::max(corpus.n_output, corpus.n_input);
    if (n_symbol) log << " per-output-symbol-perplexity(N=" << n_symbol << ")=" << fmt_base2(corpus_p.ppxper(n_symbol));
    if (corpus.n_pairs)
      log << " per-example-perplexity(N=" << corpus.n_pairs << ")=" << fmt_base2(corpus_p.ppxper(corpus.n_pairs));
  }

  // train.cc:503-678 (single start; random restarts -! are not restated: RNG dependent)
  W train() {
    W corpus_p;  // (set_composed + initial normalize already done in the constructor, train.cc:508-509)
    if (opt.max_iter == 0 || opt.max_iter == 1) {  // train.cc:520-538
      cascade.update();
      W p = estimate(corpus_p);
      history.push_back({1, corpus_p.w, p.w, 0});
      log << "Corpus ";
      print_ppx(corpus_p);
      if (opt.max_iter == 0) {
        for (auto& a : arcs.t)
          if (!a.arc->isLocked()) {
            a.scratch = a.arc->weight;
            a.arc->weight = a.counts + a.prior_counts;
          }
        cascade.distribute_counts();
      } else {
        maximize(1);
        cascade.use_counts_final(methods);
      }
      log << "\n";
      return p.ppxper(corpus.totalEmpiricalWeight);
    }
    W bestPerplexity = W::inf();
    bool using_cascade = !cascade.trivial;
    double growth = opt.learning_rate_growth_factor;
    if (using_cascade && growth != 1) growth = 1;
    bool have_good_weights = false;
    unsigned train_iter = 0;
    W lastChange = W(10);
    W lastPerplexity = W::inf();
    double learning_rate = 1;
    bool last_was_reset = false;
    for (;;) {
      const bool first_time = train_iter == 0;
      ++train_iter;
      bool cascade_counts = using_cascade && !first_time;
      if (cascade_counts)
        for (auto& a : arcs.t) a.em_weight = a.arc->weight;  // save_counts train.cc:123-125
      cascade.update();
      if (~opt.max_iter && train_iter > opt.max_iter && have_good_weights) {
        log << "Maximum number of iterations (" << opt.max_iter
            << ") reached before convergence criteria was met - greatest arc weight change was "
            << fmt_weight(lastChange) << "\n";
        break;
      }
      W p = estimate(corpus_p);
      W newPerplexity = p.ppxper(corpus.totalEmpiricalWeight);
      history.push_back({train_iter, corpus_p.w, p.w, lastChange.getReal()});
      log << "i=" << train_iter << " (rate=" << learning_rate << "): ";
      print_ppx(corpus_p);
      if (newPerplexity < bestPerplexity && (!using_cascade || cascade_counts)) {
        log << " (new best)";
        bestPerplexity = newPerplexity;
        have_good_weights = true;
        if (!cascade.trivial)
          for (auto& a : arcs.t) a.best_weight = a.em_weight;  // save_best_counts
        else
          for (auto& a : arcs.t) a.best_weight = a.arc->weight;  // save_best
      }
      W pp_ratio_scaled;
      if (first_time) {
        log << std::endl;
        log << "Initial best start point ppx=" << fmt_base2(newPerplexity) << std::endl;
        pp_ratio_scaled.setZero();
      } else {
        pp_ratio_scaled = newPerplexity.relative_perplexity_ratio(lastPerplexity);
        log << " (relative-perplexity-ratio=" << fmt_weight(pp_ratio_scaled) << ")";
        if (lastChange < W(1.)) log << ", max {d(weight)}=" << fmt_weight(lastChange);
        log << std::endl;
      }
      if (!last_was_reset) {
        if (pp_ratio_scaled >= opt.converge_perplexity_ratio) {
          if (learning_rate > 1) {
            log << "Failed to improve (relaxation rate too high); starting again at learning rate 1" << std::endl;
            learning_rate = 1;
            for (auto& a : arcs.t) a.arc->weight = a.em_weight;  // keep_em_weight
            last_was_reset = true;
            continue;
          }
          log << "Converged - per-example perplexity ratio exceeds " << fmt_weight(opt.converge_perplexity_ratio)
              << " after " << train_iter << " iterations.\n";
          if (!have_good_weights)
            log << "Because of the --train-cascade implementation, we need another iteration even though "
                   "we've converged.\n";
          else
            break;
        } else {
          if (learning_rate < 20) learning_rate *= growth;  // config.h:145
        }
      } else
        last_was_reset = false;
      lastChange = maximize(learning_rate);
      if (lastChange <= opt.converge_arc_delta && have_good_weights) {
        log << "Converged - maximum weight change less than " << fmt_weight(opt.converge_arc_delta) << " after "
            << train_iter << " iterations.\n";
        break;
      }
      lastPerplexity = newPerplexity;
    }
    log << "Setting weights to model with lowest per-example-perplexity ( = "
           "prod[modelprob(example)]^(-1/num_examples) = 2^(-log_2(p_model(corpus))/N) = "
        << fmt_base2(bestPerplexity) << std::endl;
    for (auto& a : arcs.t) a.arc->weight = a.best_weight;  // load_best
    cascade.use_counts_final(methods);
    return bestPerplexity;
  }
};

}  // namespace orc
