// forest_oracle.hpp -- CPU ORACLE for forest-em's inside-outside EM (test infrastructure, NOT the product).
//
// Plain C++17, Boost-free restatement of graehl/carmel's forest-em training path (SURVEY.md section 8
// rows a19-a23): derivation-forest reader/printer, recursive inside pass with ancestry recording,
// normalised outside pass, count accumulation (with the float near-limit spill table),
// NormalizeGroups and the overrelaxed_em driver.  Used only by tests/, smoke() and bench.py's CPU legs.
// Nothing under carmel_b200/ may include, link or execute this file.
//
// Parity status: forest-em itself ships no expected outputs (forest-em/sample/* are inputs only), so the
// pins are indirect but real (tests/test_forest_oracle.py):
//  * reader/printer: PINNED to the reference's unit-test vectors (forest.hpp:1041-1070 test_forests[]);
//  * EM numerics: PINNED to the reference's golden log through cross-program identity -- the cipher cascade
//    exported as forests (--fem-forest/--fem-norm/--fem-param, cascade.h:85-166) and trained here
//    reproduces carmel-tutorial/commands.trace:6905-6950 (2^-2245.63 ... 2^-1734.43, 22 iterations);
//  * inside scores and expected counts: checked against brute-force enumeration of every derivation on
//    the reference's sample forests and on seeded random forests.
// Float-mode rounding, the near-limit count spill and the log lines have no reference vectors: unpinned.
//
// Each function cites the reference file:line it follows (paths relative to the reference root).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace forc {

// ------------------------------------------------------------------------------------------
// logweight<Real>  (graehl/shared/weight.h:131-604; + :765-801, - :803-830, cutoff :103)
// ------------------------------------------------------------------------------------------
template <class Real>
struct LW {
  Real w;  // natural log
  static Real inf() { return std::numeric_limits<Real>::infinity(); }
  static Real much_bigger() { return sizeof(Real) == 4 ? (Real)16. : (Real)36.; }  // weight.h:103
  static Real underflow_ln() { return sizeof(Real) == 4 ? (Real)73. : (Real)82.; }  // weight.h:112
  LW() : w(-inf()) {}
  LW(Real ln, bool) : w(ln) {}
  LW(double real) { w = real > 0 ? (Real)std::log((Real)real) : -inf(); }  // setReal weight.h:295-300
  static LW ln(Real l) { return LW(l, false); }
  bool isZero() const { return !(w > -inf()); }
  bool isInfinity() const { return w == inf(); }
  bool isNearAddOneLimit() const { return w > much_bigger() - 2; }  // weight.h:282-291
  bool fitsInReal() const { return isZero() || (w < underflow_ln() && w > -underflow_ln()); }
  Real getLn() const { return w; }
  double getReal() const { return std::exp(w); }
  LW inverse() const { return LW(-w, false); }
  LW pow(Real p) const { return isZero() ? *this : LW(w * p, false); }
};
template <class R>
inline LW<R> operator*(LW<R> a, LW<R> b) { return LW<R>(a.w + b.w, false); }
template <class R>
inline LW<R> operator/(LW<R> a, LW<R> b) { return LW<R>(a.w - b.w, false); }
template <class R>
inline LW<R> operator+(LW<R> lhs, LW<R> rhs) {  // weight.h:765-801
  if (lhs.isZero()) return rhs;
  if (rhs.isZero()) return lhs;
  R diff = lhs.w - rhs.w;
  if (diff > LW<R>::much_bigger()) return lhs;
  if (diff < -LW<R>::much_bigger()) return rhs;
  if (diff < 0) return LW<R>((R)(rhs.w + log1p(std::exp(diff))), false);
  return LW<R>((R)(lhs.w + log1p(std::exp(-diff))), false);
}
template <class R>
inline LW<R> operator-(LW<R> lhs, LW<R> rhs) {  // weight.h:803-830
  if (rhs.isZero()) return lhs;
  R rdiff = rhs.w - lhs.w;
  if (rdiff >= 0) return LW<R>();
  if (rdiff < -LW<R>::much_bigger()) return lhs;
  return LW<R>((R)(lhs.w + log1p(-std::exp(rdiff))), false);
}
template <class R>
inline LW<R>& operator+=(LW<R>& a, LW<R> b) { return a = a + b; }
template <class R>
inline LW<R>& operator*=(LW<R>& a, LW<R> b) { return a = a * b; }
template <class R>
inline bool operator>(LW<R> a, LW<R> b) { return a.w > b.w; }
template <class R>
inline bool operator<(LW<R> a, LW<R> b) { return a.w < b.w; }
template <class R>
inline LW<R> absdiff(LW<R> a, LW<R> b) { return a.w > b.w ? a - b : b - a; }  // weight.h:836-855

// weight.h:463-490 print: precision 7 (float) / 15 (double); forest-em sets ALWAYS_LOG + EXP base
// (forest-em-params.cpp:75-84) unless --human-probs (NEVER_LOG).
template <class R>
inline std::string fmt_weight(LW<R> x, bool human) {
  if (x.isZero()) return "0";
  std::ostringstream o;
  o.precision(sizeof(R) > 4 ? 15 : 7);
  if (human)
    o << x.getReal();
  else
    o << "e^" << x.getLn();
  return o.str();
}
inline std::string fmt_base2(double ln) {  // weight.h:546-549 as_base(2) at stream precision 6
  std::ostringstream o;
  o.precision(6);
  o << "2^" << ln / std::log(2.);
  return o.str();
}
// weight.h:503-528 setString (forms: 0.5  e^-3  -3ln  -2log  10^-2)
inline bool parse_ln_weight(std::string const& s, double& ln_out) {
  const char* b = s.c_str();
  const char* end = b + s.size();
  char* e;
  const double ln10 = 2.30258509299404568402;
  if (b == end) return false;
  if (s.size() > 2 && b[0] == 'e' && b[1] == '^') {
    ln_out = std::strtod(b + 2, &e);
    return e == end;
  }
  if (s.size() > 3 && b[0] == '1' && b[1] == '0' && b[2] == '^') {
    ln_out = std::strtod(b + 3, &e) * ln10;
    return e == end;
  }
  double d = std::strtod(b, &e);
  if (e == b) return false;
  if (e[0] == 'l' && e[1] == 'n' && e + 2 == end) {
    ln_out = d;
    return true;
  }
  if (e[0] == 'l' && e[1] == 'o' && e[2] == 'g' && e + 3 == end) {
    ln_out = d * ln10;
    return true;
  }
  if (e != end) return false;
  ln_out = d > 0 ? std::log(d) : -std::numeric_limits<double>::infinity();
  return true;
}

// ------------------------------------------------------------------------------------------
// ForestNode / FForest  (forest-em/forest.hpp:60-76,85-817)
// ------------------------------------------------------------------------------------------
struct ForestNode {
  uint32_t next;   // index one past my last descendant (forest.hpp:62-64)
  uint32_t label;  // rule id (>0), 0 = OR, or (backref) index of the shared node
  bool backref;
  bool is_or() const { return !backref && label == 0; }
};

struct Forest {
  std::vector<ForestNode> nodes;
  size_t size() const { return nodes.size(); }

  // forest.hpp:135-242 GENIO_read.  Returns false on clean EOF before a forest starts; throws on errors.
  bool read(std::istream& in, size_t& max_ruleid) {
    nodes.clear();
    std::vector<uint32_t> open_parens;  // nodes whose `next` is set by the matching ')'
    std::unordered_map<size_t, uint32_t> backrefs;
    bool follows_paren = false, first = true;
    uint32_t n_done = 0;        // == `stop - nodes` of the reference
    std::vector<ForestNode> buf;
    auto at_stop = [&]() -> ForestNode& {
      if (buf.size() <= n_done) buf.resize(n_done + 1, ForestNode{0, 0, false});
      return buf[n_done];
    };
    while (n_done == 0 || !open_parens.empty()) {
      char c;
      if (!(in >> c)) {
        if (first) return false;
        throw std::runtime_error("Forest: unexpected end of input");
      }
      first = false;
      switch (c) {
        case '#': {
          if (follows_paren) throw std::runtime_error("Bad # following paren in Forest");
          size_t id;
          if (!(in >> id)) throw std::runtime_error("Forest: expected backreference id after #");
          if (!in.get(c)) throw std::runtime_error("Forest: unexpected end of input after #id");
          if (c == '(') {
            backrefs[id] = n_done;
          } else {
            auto it = backrefs.find(id);
            if (it == backrefs.end()) throw std::runtime_error("Forest: backreference to undefined #" + std::to_string(id));
            ForestNode& s = at_stop();
            s.label = it->second;
            s.backref = true;
            s.next = n_done + 1;
            ++n_done;
          }
          in.unget();
          break;
        }
        case '(':
          follows_paren = true;
          at_stop();
          open_parens.push_back(n_done);
          break;
        case '1': case '2': case '3': case '4': case '5': case '6': case '7': case '8': case '9': {
          in.unget();
          uint32_t rule;
          in >> rule;
          if (max_ruleid < rule) max_ruleid = rule;
          ForestNode& s = at_stop();
          s.label = rule;
          s.backref = false;
          if (!follows_paren) {
            s.next = n_done + 1;
            ++n_done;
          } else {
            follows_paren = false;
            ++n_done;
          }
          break;
        }
        case 'O': {
          char r;
          if (!in.get(r) || r != 'R') throw std::runtime_error("Forest: expected OR");
          if (!follows_paren) throw std::runtime_error("OR not following paren in Forest");
          follows_paren = false;
          ForestNode& s = at_stop();
          s.label = 0;
          s.backref = false;
          ++n_done;
          break;
        }
        case ')':
          if (open_parens.empty()) throw std::runtime_error("Forest: unbalanced )");
          buf[open_parens.back()].next = n_done;
          open_parens.pop_back();
          break;
        default:
          throw std::runtime_error(std::string("Forest: unexpected char ") + c);
      }
    }
    buf.resize(n_done);
    nodes.swap(buf);
    return true;
  }

  // forest.hpp:245-320 print (backreference ids renumbered in order of first reference)
  void print(std::ostream& o) const {
    const uint32_t n = (uint32_t)nodes.size();
    std::vector<uint32_t> ids(n, 0);
    uint32_t lastid = 0;
    for (uint32_t p = 0; p < n; ++p)
      if (nodes[p].backref && !ids[nodes[p].label]) ids[nodes[p].label] = ++lastid;
    std::vector<uint32_t> ends{n};
    bool first = true;
    for (uint32_t p = 0; p < n; ++p) {
      const uint32_t id = ids[p];
      while (p == ends.back() && ends.size() > 1) {
        o << ')';
        ends.pop_back();
      }
      if (first)
        first = false;
      else
        o << ' ';
      if (id) o << '#' << id;
      if (nodes[p].backref) {
        o << '#' << ids[nodes[p].label];
      } else {
        const uint32_t rule = nodes[p].label;
        if (nodes[p].next == p + 1) {
          if (id) o << '(';
          o << rule;
          if (id) o << ')';
        } else {
          o << '(';
          ends.push_back(nodes[p].next);
          if (rule == 0)
            o << "OR";
          else
            o << rule;
        }
      }
    }
    while (ends.size() > 1) {
      o << ')';
      ends.pop_back();
    }
  }
};

// ------------------------------------------------------------------------------------------
// NormalizeGroups  (graehl/shared/normalize.hpp:36-267)
// ------------------------------------------------------------------------------------------
struct NormGroups {
  std::vector<std::vector<size_t>> groups;
  size_t max_index = 0;
  enum { ZERO_ZEROCOUNTS = 0, SKIP_ZEROCOUNTS = 1, UNIFORM_ZEROCOUNTS = 2 };
  // normalize.hpp:58-65 read "((1 2 3) (4 5))"
  void read(std::istream& in) {
    char c;
    if (!(in >> c) || c != '(') throw std::runtime_error("normalization groups: expected (");
    for (;;) {
      if (!(in >> c)) throw std::runtime_error("normalization groups: unexpected end of input");
      if (c == ')') break;
      if (c != '(') throw std::runtime_error("normalization groups: expected ( or )");
      groups.emplace_back();
      for (;;) {
        if (!(in >> c)) throw std::runtime_error("normalization group: unexpected end of input");
        if (c == ')') break;
        in.unget();
        size_t v;
        if (!(in >> v)) throw std::runtime_error("normalization group: expected parameter index");
        groups.back().push_back(v);
        max_index = std::max(max_index, v);
      }
    }
  }
  size_t num_params() const {
    size_t n = 0;
    for (auto const& g : groups) n += g.size();
    return n;
  }
  // normalize.hpp:229-246 init_uniform: every grouped parameter <- 1 then normalised
  template <class R>
  void init_uniform(std::vector<LW<R>>& w) const {
    for (auto const& g : groups) {
      LW<R> sum;
      for (size_t i : g) {
        w[i] = LW<R>(1.);
        sum += w[i];
      }
      if (sum > LW<R>())
        for (size_t i : g) w[i] = w[i] / sum;
    }
  }
  // normalize.hpp:123-164,254-267.  Returns (maxdiff real, index).
  template <class R>
  std::pair<double, size_t> normalize(std::vector<LW<R>> const& src, std::vector<LW<R>>& dst, LW<R> add_k, int zerocounts,
                                      std::ostream* log) const {
    LW<R> maxdiff;
    size_t maxdiff_index = 0;
    size_t gi = 0;
    for (auto const& g : groups) {
      ++gi;
      LW<R> sum;
      for (size_t j : g) sum += src[j];
      auto dodiff = [&](LW<R> d, LW<R> w, size_t j) {
        LW<R> diff = absdiff(d, w);
        if (maxdiff < diff) {
          maxdiff_index = j;
          maxdiff = diff;
        }
      };
      if (sum > LW<R>()) {
        sum += add_k;
        for (size_t j : g) {
          LW<R> prev = dst[j];
          LW<R> w = src[j];  // src and dst may alias
          dst[j] = w / sum;
          dodiff(dst[j], prev, j);
        }
      } else {
        if (log)
          *log << "Zero counts for normalization group #" << gi << " with first parameter " << g.front() << " (one of "
               << g.size() << " parameters)";
        if (zerocounts != SKIP_ZEROCOUNTS) {
          LW<R> setto;
          if (zerocounts == UNIFORM_ZEROCOUNTS) {
            setto = LW<R>(1. / (double)g.size());
            if (log) *log << " - setting to uniform probability " << fmt_weight(setto, false) << std::endl;
          } else if (log)
            *log << " - setting to zero probability." << std::endl;
          for (size_t j : g) {
            dodiff(dst[j], setto, j);
            dst[j] = setto;
          }
        }
      }
    }
    return {maxdiff.getReal(), maxdiff_index};
  }
};

// ------------------------------------------------------------------------------------------
// FForests  (forest-em/forest-em.hpp:48-692) + inside/outside of FForest (forest.hpp:326-491,636-697)
// ------------------------------------------------------------------------------------------
struct ForestOpts {
  unsigned max_iter = 1000;             // -i  (forest-em-params.hpp:185)
  double converge_ratio = 1. / 65536;   // -e  (:186)
  double converge_delta = 0;            // -d  (:187)
  double prior_counts = 0;              // -p
  double add_k_smoothing = 0;           // -k
  bool zero_zerocounts = false;         // -z
  bool initial_1_params = false;        // -u
  bool normalize_initial = false;       // -N
  bool human_probs = false;             // -H
  unsigned log_level = 1;
};

template <class Real>
struct Forests {
  typedef LW<Real> W;
  std::vector<Forest> forests;
  NormGroups norm_groups;
  std::vector<W> rule_weights, counts;
  bool have_init_params = false;
  size_t max_forest_ruleid = 0, max_nodes = 0, n_nodes = 0, rulespace = 0;
  ForestOpts opt;
  // per-run state
  std::vector<W> inside, outside;
  struct Ancestry {
    uint32_t parent, child;
  };
  std::vector<Ancestry> outside_order;
  std::unordered_map<unsigned, W> overflows;  // forest.hpp:352 count_overflows
  unsigned n_overflows = 0;
  double total_logprob = 0;
  size_t n_zeroprob = 0, forest_no = 0;
  bool firsttime = true;
  unsigned iteration = 0;
  std::vector<double> last_inside;  // ln inside[0] of every forest in the last estimate (for tests / -S)
  struct Iter {
    unsigned i;
    double alp;
    double max_delta;
    size_t max_index;
    size_t n;
  };
  std::vector<Iter> history;

  // forest-em.hpp:228-250 read_params: weight #k is parameter k (1-based); optional surrounding parens
  void read_params(std::istream& in) {
    rule_weights.clear();
    rule_weights.push_back(W());
    std::string tok;
    while (in >> tok) {
      if (tok == "(" || tok == ")") continue;
      if (tok[0] == '(') tok = tok.substr(1);
      if (!tok.empty() && tok.back() == ')') tok.pop_back();
      if (tok.empty()) continue;
      double ln;
      if (!parse_ln_weight(tok, ln)) throw std::runtime_error("couldn't read rule weights: bad weight " + tok);
      rule_weights.push_back(W::ln((Real)ln));
    }
    have_init_params = true;
  }
  void read_norm_groups(std::istream& in) {  // forest-em.hpp:133-149
    norm_groups.read(in);
    if (have_init_params && rule_weights.size() <= norm_groups.max_index)
      throw std::runtime_error("Initial rule weights file not big enough - normalization used rule (" +
                               std::to_string(norm_groups.max_index) + " expected)");
  }
  void read_forests(std::istream& in) {  // forest-em.hpp:150-168
    for (;;) {
      Forest f;
      if (!f.read(in, max_forest_ruleid)) break;
      max_nodes = std::max(max_nodes, f.size());
      n_nodes += f.size();
      forests.push_back(std::move(f));
    }
  }
  // forest-em.hpp:335-381 prepare + :282-306 init_rule_weights
  void prepare() {
    rulespace = std::max(max_forest_ruleid, norm_groups.max_index) + 1;
    inside.assign(max_nodes, W());
    outside.assign(max_nodes, W());
    if (have_init_params) {
      if (rulespace > rule_weights.size()) throw std::runtime_error("Initial params file wasn't large enough for forests/norms.");
    } else if (opt.initial_1_params) {
      rule_weights.assign(rulespace, W(1.));
    } else {
      rule_weights.assign(rulespace, W());
      norm_groups.init_uniform(rule_weights);
    }
    counts.assign(rulespace, W());  // forest-em.hpp:362 counts.alloc(rulespace)
    firsttime = true;
    iteration = 0;
  }
  void normalize_params() {  // forest-em.hpp:618-621
    norm_groups.normalize(rule_weights, rule_weights, W(), NormGroups::UNIFORM_ZEROCOUNTS, nullptr);
  }

  // forest.hpp:636-697 inside_rec (+ ancestry recording)
  void inside_rec(Forest const& f, uint32_t b) {
    const uint32_t e = f.nodes[b].next;
    const uint32_t i = b;
    ForestNode const& nd = f.nodes[b];
    if (nd.backref) {
      inside[i] = inside[nd.label];
      return;
    }
    const uint32_t parent = b;
    if (nd.label == 0) {  // OR
      ++b;
      uint32_t n = f.nodes[b].next;
      inside_rec(f, b);
      inside[i] = inside[i + 1];
      for (b = n; b < e; b = n) {
        n = f.nodes[b].next;
        inside_rec(f, b);
        inside[i] += inside[b];
      }
    } else {  // AND
      inside[i] = rule_weights[nd.label];
      ++b;
      uint32_t n;
      for (; b < e; b = n) {
        n = f.nodes[b].next;
        inside_rec(f, b);
        inside[i] *= inside[b];
      }
    }
    b = parent + 1;
    if (b != e) do {
        ForestNode const& c = f.nodes[b];
        outside_order.push_back(Ancestry{parent, c.backref ? c.label : b});
        b = c.next;
      } while (b != e);
  }
  // forest.hpp:507-574 viterbi_rec: as inside_rec with max for the OR nodes; best[i] = the chosen child of OR node i
  // (the first child that attains the maximum: a later child replaces the choice only when strictly better)
  std::vector<uint32_t> vit_best;
  void viterbi_rec(Forest const& f, uint32_t b) {
    const uint32_t e = f.nodes[b].next, i = b;
    ForestNode const& nd = f.nodes[b];
    if (nd.backref) {
      inside[i] = inside[nd.label];
      return;
    }
    if (nd.label == 0) {  // OR
      ++b;
      uint32_t n = f.nodes[b].next;
      viterbi_rec(f, b);
      inside[i] = inside[i + 1];
      vit_best[i] = b;
      for (b = n; b < e; b = n) {
        n = f.nodes[b].next;
        viterbi_rec(f, b);
        if (inside[i] < inside[b]) {
          inside[i] = inside[b];
          vit_best[i] = b;
        }
      }
    } else {  // AND
      inside[i] = rule_weights[nd.label];
      ++b;
      uint32_t n;
      for (; b < e; b = n) {
        n = f.nodes[b].next;
        viterbi_rec(f, b);
        inside[i] *= inside[b];
      }
    }
  }
  // forest.hpp:590-631 write_viterbi_rec
  void write_viterbi_rec(std::ostream& o, Forest const& f, uint32_t b) const {
    ForestNode const& nd = f.nodes[b];
    if (nd.backref) return write_viterbi_rec(o, f, nd.label);
    if (nd.label == 0) return write_viterbi_rec(o, f, vit_best[b]);
    const uint32_t e = nd.next;
    if (b + 1 == e) {
      o << nd.label;
      return;
    }
    o << '(' << nd.label;
    uint32_t n;
    for (++b; b < e; b = n) {
      n = f.nodes[b].next;
      o << ' ';
      write_viterbi_rec(o, f, b);
    }
    o << ')';
  }
  // forest-em.hpp:535-550 (final viterbi decoding, -v): "best/sum=pct% tree" per forest (forest.hpp:581-585)
  void write_viterbi(std::ostream& o, Forest const& f, bool human) {
    inside_rec(f, 0);
    outside_order.clear();
    const W sum = inside[0];
    vit_best.assign(f.size(), 0);
    viterbi_rec(f, 0);
    o << fmt_weight(inside[0], human) << '/' << fmt_weight(sum, human) << '=' << 100 * (inside[0] / sum).getReal() << "% ";
    write_viterbi_rec(o, f, 0);
    o << '\n';
  }
  // ------------------------------------------------------------------------------------------------------------
  // Gibbs sampling over forests (forest-em --crp): FForests::run_gibbs / to_gibbs / resample_block / from_gibbs
  // (forest-em/forest-em.hpp:694-797), FForest::compute_inside(W) + choose_random (forest/forest.hpp:726-816),
  // gibbs_base (graehl/shared/gibbs.hpp:582-623,712-792,803-877), delta_sum (delta_sum.hpp:49-106).
  // PARITY UNPINNED by the reference (forest-em ships no expected outputs and draws from boost's generator): sampled
  // derivations are defined on injected uniforms u(seed, sweep, forest, draw), one per OR node visited, in visit order.
  // ------------------------------------------------------------------------------------------------------------
  struct GibbsOpts {
    unsigned iter = 0, burnin = 0;
    double alpha = .1;  // --const-alpha (gibbs_opts.hpp:229)
    bool uniformp0 = false, final_counts = false, exclude_prior = false, sample_prob = false;
    double high_temp = 1, low_temp = 1;
    uint64_t seed = 1;
    double n_sym = 0;  // --n-symbols; default: total forest nodes (forest-em.hpp:733)
  };
  struct GDelta {  // delta_sum.hpp:49-106
    double x = 0, tmax = 0, s = 0;
    void clear(double x0) { x = x0, s = tmax = 0; }
    void add_delta(double d, double t) {
      const double moret = t - tmax;
      if (moret > 0) {
        tmax = t;
        s += moret * x;
      } else if (moret < 0)
        s += d * (-moret);
      x += d;
    }
    void extend(double t) {
      const double moret = t - tmax;
      tmax = t;
      s += x * moret;
    }
  };
  static constexpr unsigned G_NONORM = 0xFFFFFFFFu;
  std::vector<double> g_prior;       // pseudo-count, or the fixed probability of a parameter without a group
  std::vector<unsigned> g_norm;      // normalisation group (1-based as visit_norm_param numbers them) or G_NONORM
  std::vector<GDelta> g_count;
  std::vector<double> g_normsum;
  std::vector<std::vector<unsigned>> g_sample;  // per forest: rule ids in record order
  std::vector<double> g_iter_ln_prob;
  static uint64_t g_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
  }
  static double g_uniform(uint64_t seed, uint32_t sweep, uint32_t block, uint32_t draw) {  // as gibbs_oracle.hpp
    uint64_t h = g_mix64(seed ^ g_mix64(((uint64_t)sweep << 32) | block));
    h = g_mix64(h + draw);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
  }
  double g_proposal(unsigned id) const {  // gibbs.hpp:154-157
    return g_norm[id] != G_NONORM ? g_count[id].x / g_normsum[g_norm[id]] : g_prior[id];
  }
  // forest.hpp:769-816 compute_inside(W): inside with the proposal probabilities
  void g_inside_rec(Forest const& f, uint32_t b) {
    const uint32_t e = f.nodes[b].next, i = b;
    ForestNode const& nd = f.nodes[b];
    if (nd.backref) {
      inside[i] = inside[nd.label];
      return;
    }
    if (nd.label == 0) {
      ++b;
      uint32_t n = f.nodes[b].next;
      g_inside_rec(f, b);
      inside[i] = inside[i + 1];
      for (b = n; b < e; b = n) {
        n = f.nodes[b].next;
        g_inside_rec(f, b);
        inside[i] += inside[b];
      }
    } else {
      inside[i] = W((Real)g_proposal(nd.label));
      ++b;
      uint32_t n;
      for (; b < e; b = n) {
        n = f.nodes[b].next;
        g_inside_rec(f, b);
        inside[i] *= inside[b];
      }
    }
  }
  // forest.hpp:726-758 choose_random.  NB the reference follows a back reference with the default power = 1
  // (`choose_random(l.pointer(), v)`), so annealing stops below shared sub-forests; reproduced.
  void g_choose(Forest const& f, uint32_t b, double power, GibbsOpts const& o, uint32_t sweep, uint32_t block, uint32_t& draw,
                std::vector<unsigned>& out) {
    ForestNode const& nd = f.nodes[b];
    const uint32_t e = nd.next;
    if (nd.backref) return g_choose(f, nd.label, 1., o, sweep, block, draw, out);
    if (nd.label == 0) {
      W norm;  // zero
      for (uint32_t i = b + 1; i != e; i = f.nodes[i].next) norm += inside[i].pow((Real)power);
      uint32_t i = b + 1;
      double choice = g_uniform(o.seed, sweep, block, draw++);
      for (;;) {
        choice -= (inside[i].pow((Real)power) / norm).getReal();
        if (choice < 0) break;
        const uint32_t n = f.nodes[i].next;
        if (n == e) break;
        i = n;
      }
      g_choose(f, i, power, o, sweep, block, draw, out);
    } else {
      out.push_back(nd.label);
      for (uint32_t c = b + 1; c < e; c = f.nodes[c].next) g_choose(f, c, power, o, sweep, block, draw, out);
    }
  }
  void g_addc(std::vector<unsigned> const& ids, double d, double time) {  // gibbs.hpp:769-792
    for (unsigned id : ids)
      if (g_norm[id] != G_NONORM) {
        g_normsum[g_norm[id]] += d;
        g_count[id].add_delta(d, time);
      }
  }
  void run_gibbs(GibbsOpts o, std::ostream& log) {
    // to_gibbs (forest-em.hpp:731-742): normalise, then one parameter per rule (visit_norm_param, normalize.hpp:194-210)
    normalize_params();
    g_prior.assign(rulespace, 0.);
    g_norm.assign(rulespace, G_NONORM);
    unsigned normi = 0;
    for (auto const& g : norm_groups.groups) {
      ++normi;
      for (size_t p : g) {
        g_norm[p] = normi;
        g_prior[p] = o.uniformp0 ? o.alpha : o.alpha * (double)rule_weights[p].getReal() * (double)g.size();
      }
    }
    for (size_t p = 0; p < rulespace; ++p)
      if (g_norm[p] == G_NONORM) g_prior[p] = (double)rule_weights[p].getReal();
    const unsigned nnorm = normi + 1;
    if (!o.n_sym) o.n_sym = (double)n_nodes;
    // run (gibbs.hpp:803-828)
    g_count.assign(rulespace, GDelta());
    g_normsum.assign(nnorm, 0.);
    for (size_t p = 0; p < rulespace; ++p)
      if (g_norm[p] != G_NONORM) {
        g_normsum[g_norm[p]] += g_prior[p];
        g_count[p].clear(g_prior[p]);
      }
    g_sample.assign(forests.size(), {});
    g_iter_ln_prob.clear();
    for (unsigned iter = 0; iter <= o.iter; ++iter) {
      double time = (double)iter - (double)o.burnin;
      if (time < 0 || iter == 0) time = 0;
      double temperature = o.high_temp;
      if (o.iter > 0 && o.high_temp != o.low_temp)
        temperature = o.high_temp + (o.low_temp - o.high_temp) * std::min(1.0, (double)iter / o.iter);
      const double power = temperature > 0 ? 1. / temperature : 1;
      std::vector<double> ccount(rulespace, 0.), csum(nnorm, 0.);  // cache reset (gibbs.hpp:656-667,700-705)
      for (size_t p = 0; p < rulespace; ++p)
        if (g_norm[p] != G_NONORM) csum[g_norm[p]] += (ccount[p] = g_prior[p]);
      double ln_p = 0;
      for (uint32_t b = 0; b < forests.size(); ++b) {
        Forest const& f = forests[b];
        g_addc(g_sample[b], -1., time);
        g_sample[b].clear();
        g_inside_rec(f, 0);
        uint32_t draw = 0;
        g_choose(f, 0, power, o, iter, b, draw, g_sample[b]);
        if (o.sample_prob) {  // as the carmel oracle: scored with the new sample's counts back in
          g_addc(g_sample[b], 1., time);
          for (unsigned id : g_sample[b]) ln_p += std::log(g_proposal(id));
          continue;
        }
        for (unsigned id : g_sample[b])
          ln_p += std::log(g_norm[id] != G_NONORM ? ccount[id]++ / csum[g_norm[id]]++ : g_prior[id]);
        g_addc(g_sample[b], 1., time);
      }
      g_iter_ln_prob.push_back(ln_p);
      const double l2 = 1. / std::log(2.);
      log << "Gibbs i=" << iter << (o.sample_prob ? " sample prob=" : " cache-model prob=") << "2^" << fmt_g6(ln_p * l2);
      if (o.n_sym) log << " per-point-ppx(N=" << o.n_sym << ")=2^" << fmt_g6(-ln_p * l2 / o.n_sym);
      log << " per-block-ppx(N=" << forests.size() << ")=2^" << fmt_g6(-ln_p * l2 / (double)forests.size()) << "\n";
    }
    // finalize_cumulative_counts (gibbs.hpp:626-644) + from_gibbs (forest-em.hpp:743-750)
    if (!(o.final_counts && !o.exclude_prior)) {
      const double tmax1 = ((double)o.iter - (double)o.burnin) + 1;
      for (size_t p = 0; p < rulespace; ++p) {
        if (g_norm[p] == G_NONORM) continue;
        if (o.exclude_prior) {
          g_count[p].s += -g_prior[p] * g_count[p].tmax;
          g_count[p].x += -g_prior[p];
        }
        if (!o.final_counts) {
          g_count[p].extend(tmax1);
          g_count[p].x = g_count[p].s;
        }
      }
      g_normsum.assign(nnorm, 0.);
      for (size_t p = 0; p < rulespace; ++p)
        if (g_norm[p] != G_NONORM) g_normsum[g_norm[p]] += g_count[p].x;
    }
    for (size_t p = 0; p < rulespace; ++p) {
      double fp = g_prior[p];
      if (g_norm[p] != G_NONORM) fp = g_count[p].x > 0 ? g_count[p].x / g_normsum[g_norm[p]] : 0.;
      rule_weights[p] = W((Real)fp);
    }
  }
  static std::string fmt_g6(double v) {
    std::ostringstream o;
    o << std::setprecision(6) << v;
    return o.str();
  }
  // forest.hpp:439-491 compute_norm_outside
  bool compute_norm_outside(Forest const& f) {
    if (!(inside[0] > W())) {
      std::cerr << "\nCan't collect counts when inside[0] == 0!\n";
      return false;
    }
    const size_t n = f.size();
    outside[0] = inside[0].inverse();
    for (size_t k = 1; k < n; ++k) outside[k] = W();
    for (size_t k = outside_order.size(); k > 0;) {
      --k;
      const uint32_t p = outside_order[k].parent, c = outside_order[k].child;
      if (f.nodes[p].label == 0) {
        outside[c] += outside[p];
      } else if (!inside[p].isZero()) {
        outside[c] += outside[p] * inside[p] / inside[c];
      }
    }
    return true;
  }
  // forest.hpp:353-395 accumulate_counts::operator()
  void accumulate(unsigned rule, W in, W no) {
    if (counts[rule].isNearAddOneLimit()) {
      overflows[rule] += counts[rule];
      ++n_overflows;
      counts[rule] = in * no;
    } else
      counts[rule] += in * no;
  }
  // forest.hpp:426-438 visit_inside_norm_outside
  void visit_inside_norm_outside(Forest const& f) {
    for (uint32_t i = 0, e = (uint32_t)f.size(); i != e; ++i) {
      ForestNode const& nd = f.nodes[i];
      if (!nd.backref && nd.label != 0) accumulate(nd.label, inside[i], outside[i]);
    }
  }
  // forest-em.hpp:511-551 operator()(Forest&)
  void visit_forest(Forest const& f, bool collect, bool first_time, std::ostream& log) {
    ++forest_no;
    outside_order.clear();
    inside_rec(f, 0);
    W sum = inside[0];
    if (collect && compute_norm_outside(f)) visit_inside_norm_outside(f);
    last_inside.push_back((double)sum.getLn());
    if (inside[0].isZero()) {
      if (first_time) log << "Warning: 0 probability for forest #" << forest_no << std::endl;
      ++n_zeroprob;
    } else
      total_logprob += (double)sum.getLn();
  }
  double size() const { return (double)forests.size(); }
  // forest-em.hpp:556-572 estimate (+ begin_visit :446-458, end_visit :459-468)
  double estimate(bool first_time, std::ostream& log) {
    W weighted_prior = W(opt.prior_counts) * W((double)forests.size());
    std::fill(counts.begin(), counts.end(), weighted_prior);
    total_logprob = 0;
    n_overflows = 0;
    forest_no = 0;
    n_zeroprob = 0;
    last_inside.clear();
    for (auto const& f : forests) visit_forest(f, true, first_time, log);
    for (auto const& kv : overflows) counts[kv.first] += kv.second;  // finish_counts (forest.hpp:396-407)
    overflows.clear();
    const size_t N = forest_no - n_zeroprob;
    log << "\nN=" << N << ' ';
    if (n_zeroprob) log << '(' << n_zeroprob << " 0 prob removed) ";
    return total_logprob / (double)(forest_no - n_zeroprob);
  }
  // forest-em.hpp:626-655 maximize
  std::pair<double, size_t> maximize(std::ostream& log) {
    std::ostream* logs = nullptr;
    if (firsttime) {
      firsttime = false;
      if (opt.log_level > 1) logs = &log;
    }
    int z = opt.zero_zerocounts ? NormGroups::ZERO_ZEROCOUNTS : NormGroups::UNIFORM_ZEROCOUNTS;
    auto r = norm_groups.normalize(counts, rule_weights, W(opt.add_k_smoothing), z, logs);
    ++iteration;
    return r;
  }
  static void print_alp(std::ostream& logs, double N, double alp) {  // em.hpp:101-105, weight.h:331-337
    const double ln_prob = alp * N;
    logs << "probability=" << fmt_base2(ln_prob);
    if (N > 0) logs << " per-example-perplexity(N=" << N << ")=" << fmt_base2(-ln_prob / N);
  }
  static std::string fmt_delta(std::pair<double, size_t> const& p) {  // em.hpp:60-67
    std::ostringstream o;
    if (p.first > 0)
      o << "delta_weight[" << p.second << "]=" << p.first;
    else
      o << "unchanged";
    return o.str();
  }
  // graehl/shared/em.hpp:107-216 overrelaxed_em with learning_rate_growth_factor = 1 and no random restarts
  // (forest-em-params.cpp:113).  Returns the best average log prob.
  double train(std::ostream& logs) {
    double best_alp = -HUGE_VAL;
    if (opt.max_iter == 0) return best_alp;
    const double rel_eps = opt.converge_ratio;
    bool very_first_time = true;
    const double N = size();
    unsigned train_iter = 0;
    std::pair<double, size_t> max_delta_param{0, 0};
    double last_alp = -HUGE_VAL;
    bool first_time = true;
    for (;;) {
      ++train_iter;
      if (train_iter > opt.max_iter) {
        logs << "Maximum number of iterations (" << opt.max_iter
             << ") reached before convergence criteria was met - greatest param weight change was " << fmt_delta(max_delta_param)
             << "\n";
        break;
      }
      double new_alp = estimate(very_first_time, logs);
      logs << "i=" << train_iter << ": ";
      print_alp(logs, N, new_alp);
      if (new_alp > best_alp || very_first_time) {
        logs << " (new best)";
        best_alp = new_alp;
      }
      very_first_time = false;
      double dpp = new_alp - last_alp;
      double last_abs = std::fabs(last_alp);
      double rel_dpp = dpp;
      if (last_abs < 1e-10) last_abs = 1e-10;  // LOGPROB_EPSILON
      rel_dpp /= last_abs;
      if (first_time) {
        rel_dpp = HUGE_VAL;
        logs << std::endl;
        first_time = false;
      } else
        logs << " (relative-d-avg-logprob=" << rel_dpp << "), max " << fmt_delta(max_delta_param) << std::endl;
      history.push_back({train_iter, new_alp, max_delta_param.first, max_delta_param.second, forest_no - n_zeroprob});
      if (rel_dpp < rel_eps) {
        logs << "\nConverged - relative per-example avg-logprob change less than " << rel_eps << " after " << train_iter
             << " iterations.\n";
        break;
      }
      max_delta_param = maximize(logs);
      if (max_delta_param.first <= opt.converge_delta) {
        logs << "\nConverged - all weights changed no more than " << opt.converge_delta << " after " << train_iter
             << " iterations.\n";
        break;
      }
      last_alp = new_alp;
    }
    logs << "\nSetting weights to model with best ";
    print_alp(logs, N, best_alp);
    logs << std::endl;
    return best_alp;
  }
  // forest-em.hpp:190-201 write_params / write_counts (io.hpp:326-355 multiline, no parens)
  void write_range(std::ostream& out, std::vector<W> const& v) const {
    for (size_t i = 1; i < v.size(); ++i) out << ' ' << fmt_weight(v[i], opt.human_probs) << "\n";
    out << std::endl;
  }
};

}  // namespace forc
