// Contraction-tree programs: the device-side replacement for OMEinsum's NestedEinsum /
// SlicedEinsum recursion (src/peps.jl:109,321), for `expect` over the terms of an `Add`
// (src/peps.jl:469-477) and for the TensorAD reverse pass over those contractions
// (src/TensorAD/TensorAD.jl:148-164 with the einsum rule of src/TensorAD/rules.jl:1-8).
//
// A tree (+ optional Hamiltonian terms, each replacing a few leaves) is compiled on the host
// into a flat program of contraction ops over "values":
//   * forward: one op per binary node for the unmodified network ("base" variant); terms are
//     handled by a dynamic programme over the tree: a node keeps one variant per distinct
//     partial replacement pattern that reaches it plus one running sum ("sigma") of all terms
//     already completed below it -- completed terms are linear in everything above, so they
//     share one chain instead of one full contraction per term (the reference does K full
//     contractions, src/peps.jl:470-476);
//   * backward: source-transformed reverse mode, dA = dC * B^H and dB = A^H * dC per op, with
//     the conjugations fused into operand loads (einsum_grad's two conj copies disappear);
//   * memory: intermediates get static offsets in one workspace arena from exact lifetimes.
//     A tree may be cut into checkpointed segments (tne.h): values crossing a segment boundary
//     are kept, the reverse pass walks the segments backwards and re-runs each segment's forward
//     ops from its checkpoints, so one segment's tape is alive at a time; inside a segment its
//     own sliced labels are looped over (every value carrying the label is 1/size as large;
//     checkpoints are read and written through strided views or accumulated);
//   * slicing: globally sliced labels are looped over (sharded by rank), leaves become strided views.
#include <algorithm>
#include <functional>
#include <set>

#include "tne_internal.h"

namespace {

struct Val {
    std::vector<int32_t> labels;  // memory order, fastest first
    std::vector<int64_t> dims;
    std::vector<int64_t> strides;  // empty: contiguous column-major
    int64_t nelem = 1;
    cplx* ext = nullptr;   // external storage (leaf, replacement, output); else workspace
    bool conj = false;     // conjugate while loading (leaf flags)
    bool needs_grad = false;
    int grad = -1;
    int64_t offset = -1;   // workspace byte offset
    int leaf = -1;         // leaf index for sliced views
    int64_t slice_base = 0;  // element offset added per slice (filled at run time)
    bool persistent = false;  // touched by more than one segment: lives in the checkpoint arena
    // where `ext` comes from, so that a cached program can be re-bound to the buffers of a later call:
    // 0 none, 1 leaf[bind_index], 2 replacement tensor (canonical index), 3 value output, 4 seed, 5 gradient[bind_index]
    int bind_kind = 0, bind_index = 0;
};

enum OpKind { OP_CONTRACT = 0, OP_AXPY = 1, OP_ZERO = 2 };

struct Op {
    int kind = OP_CONTRACT;
    int out = -1, a = -1, b = -1;
    bool conj_a = false, conj_b = false;
    bool acc = false;
    double are = 1.0, aim = 0.0;
    bool recompute = false;
    bool backward = false;
    int node = 0;          // tree node (post-order index) the op belongs to
    int seg = 0;
};

typedef std::vector<std::pair<int32_t, int64_t>> Key;  // sorted (leaf, replacement handle)

struct NodeInfo {
    std::vector<int> leaves;       // sorted leaf ids under this node
    std::map<Key, int> var;        // pattern -> value id (empty key = base)
    int sigma = -1;
    std::vector<int32_t> out_labels;
};

struct Program {
    tne_ctx* ctx = nullptr;
    std::vector<Val> vals;
    std::vector<Op> fwd, bwd, all;
    std::vector<bool> written;
    int root_val = -1;
    int64_t ws_bytes = 0;
    double flops = 0, bytes = 0, re_flops = 0;
    int64_t n_contract = 0;

    int new_val(const std::vector<int32_t>& labels, const std::vector<int64_t>& dims) {
        Val v;
        v.labels = labels;
        v.dims = dims;
        v.nelem = 1;
        for (auto d : dims) v.nelem *= d;
        vals.push_back(v);
        written.push_back(false);
        return (int)vals.size() - 1;
    }
    void emit(std::vector<Op>& list, Op op) {
        op.acc = written[op.out];
        written[op.out] = true;
        list.push_back(op);
    }
};

int find_in(const std::vector<int32_t>& v, int32_t l) {
    for (size_t i = 0; i < v.size(); ++i) if (v[i] == l) return (int)i;
    return -1;
}

// layout of an intermediate: A's surviving legs in A's order, then B's surviving legs in B's order
void out_layout(const Val& a, const Val& b, const std::vector<int32_t>& keep, std::vector<int32_t>& labels,
                std::vector<int64_t>& dims) {
    labels.clear(); dims.clear();
    for (size_t i = 0; i < a.labels.size(); ++i)
        if (find_in(keep, a.labels[i]) >= 0) { labels.push_back(a.labels[i]); dims.push_back(a.dims[i]); }
    for (size_t i = 0; i < b.labels.size(); ++i)
        if (find_in(keep, b.labels[i]) >= 0 && find_in(labels, b.labels[i]) < 0) { labels.push_back(b.labels[i]); dims.push_back(b.dims[i]); }
}

// ---------------------------------------------------------------------------------------
// tree parsing and the term dynamic programme
// ---------------------------------------------------------------------------------------
struct Builder {
    tne_ctx* ctx;
    Program& P;
    const tne_tree* tree;
    std::vector<NodeInfo> nodes;           // leaves then internal nodes
    std::vector<std::vector<int32_t>> leaf_labels;
    std::vector<int32_t> sliced;
    int n_terms = 0;
    const tne_term* terms = nullptr;
    int cur_node = 0;                      // node the ops being emitted belong to
    std::map<int64_t, int> repl_id;        // replacement handle -> canonical index (order of first appearance over the terms)

    Builder(tne_ctx* c, Program& p, const tne_tree* t) : ctx(c), P(p), tree(t) {}

    Key restrict_key(int k, const std::vector<int>& leaves) const {
        Key key;
        for (int r = 0; r < terms[k].n_repl; ++r)
            if (std::binary_search(leaves.begin(), leaves.end(), (int)terms[k].leaf[r]))
                key.emplace_back(terms[k].leaf[r], terms[k].repl[r]);
        std::sort(key.begin(), key.end());
        return key;
    }

    int contract_vals(int a, int b, const std::vector<int32_t>& keep, int out /* -1: new */, double are, double aim,
                      const std::vector<int32_t>* forced_order = nullptr) {
        if (out < 0) {
            std::vector<int32_t> labels;
            std::vector<int64_t> dims;
            if (forced_order) {
                for (auto l : *forced_order) {
                    int ia = find_in(P.vals[a].labels, l), ib = find_in(P.vals[b].labels, l);
                    labels.push_back(l);
                    dims.push_back(ia >= 0 ? P.vals[a].dims[ia] : P.vals[b].dims[ib]);
                }
            } else {
                out_layout(P.vals[a], P.vals[b], keep, labels, dims);
            }
            out = P.new_val(labels, dims);
        }
        Op op;
        op.kind = OP_CONTRACT;
        op.out = out; op.a = a; op.b = b;
        op.conj_a = P.vals[a].conj; op.conj_b = P.vals[b].conj;
        op.are = are; op.aim = aim;
        op.node = cur_node;
        P.emit(P.fwd, op);
        return out;
    }
};

int build_forward(Builder& B, const tne_buf* leaves, const int32_t* leaf_conj, bool root_is_output, int* root_out) {
    tne_ctx* ctx = B.ctx;
    Program& P = B.P;
    const tne_tree* T = B.tree;
    const int nl = T->n_leaves, nn = T->n_nodes;
    if (nl <= 0 || nn != nl - 1) return tne_fail(ctx, TNE_ERR_INVALID, "tree must be binary: %d leaves need %d nodes (got %d)", nl, nl - 1, nn);
    B.sliced.assign(T->sliced_labels, T->sliced_labels + T->n_sliced);
    B.nodes.resize(nl + nn);
    // leaves
    int pos = 0;
    for (int i = 0; i < nl; ++i) {
        Buf* b = buf_get(ctx, leaves[i]);
        if (!b) return TNE_ERR_INVALID;
        int rk = T->leaf_rank[i];
        if ((int)b->dims.size() != rk) return tne_fail(ctx, TNE_ERR_INVALID, "leaf %d has rank %d but %d labels", i, (int)b->dims.size(), rk);
        std::vector<int32_t> labels, full(T->leaf_labels + pos, T->leaf_labels + pos + rk);
        std::vector<int64_t> dims, strides;
        int64_t st = 1;
        for (int q = 0; q < rk; ++q) {
            if (find_in(B.sliced, full[q]) < 0) { labels.push_back(full[q]); dims.push_back(b->dims[q]); strides.push_back(st); }
            st *= b->dims[q];
        }
        B.leaf_labels.push_back(full);
        int v = P.new_val(labels, dims);
        P.vals[v].strides = strides;
        P.vals[v].ext = b->ptr();
        P.vals[v].bind_kind = 1; P.vals[v].bind_index = i;
        P.vals[v].conj = leaf_conj ? leaf_conj[i] != 0 : false;
        P.vals[v].leaf = i;
        P.written[v] = true;
        B.nodes[i].leaves = {i};
        B.nodes[i].var[Key()] = v;
        pos += rk;
    }
    // replacement leaves and leaf-level sigma
    for (int k = 0; k < B.n_terms; ++k) {
        const tne_term& t = B.terms[k];
        if (t.n_repl < 0 || t.n_repl > 4) return tne_fail(ctx, TNE_ERR_INVALID, "term %d: n_repl must be 0..4", k);
        for (int r = 0; r < t.n_repl; ++r) {
            int lf = t.leaf[r];
            if (lf < 0 || lf >= nl) return tne_fail(ctx, TNE_ERR_INVALID, "term %d replaces unknown leaf %d", k, lf);
            Key key = {{(int32_t)lf, t.repl[r]}};
            if (B.nodes[lf].var.count(key)) continue;
            Buf* b = buf_get(ctx, t.repl[r]);
            if (!b) return TNE_ERR_INVALID;
            const Val base = P.vals[B.nodes[lf].var[Key()]];
            Buf* lb = buf_get(ctx, leaves[lf]);
            if (b->dims != lb->dims) return tne_fail(ctx, TNE_ERR_INVALID, "term %d: replacement for leaf %d has a different shape", k, lf);
            int v = P.new_val(base.labels, base.dims);
            P.vals[v].strides = base.strides;
            P.vals[v].ext = b->ptr();
            P.vals[v].bind_kind = 2; P.vals[v].bind_index = B.repl_id.at(t.repl[r]);
            P.vals[v].conj = base.conj;
            P.vals[v].leaf = lf;
            P.written[v] = true;
            B.nodes[lf].var[key] = v;
        }
    }
    std::vector<int> parent_of_leaf(nl, 0);
    for (int j = 0; j < nn; ++j)
        for (int q = 0; q < 2; ++q) { int c = T->node_children[2 * j + q]; if (c >= 0 && c < nl) parent_of_leaf[c] = j; }
    for (int k = 0; k < B.n_terms; ++k) {
        const tne_term& t = B.terms[k];
        if (t.n_repl != 1) continue;
        int lf = t.leaf[0];
        NodeInfo& nd = B.nodes[lf];
        const Val base = P.vals[nd.var[Key()]];
        if (nd.sigma < 0) nd.sigma = P.new_val(base.labels, base.dims);  // dense workspace copy, conj already applied
        Op op;
        op.kind = OP_AXPY;
        op.out = nd.sigma;
        op.a = nd.var[Key{{(int32_t)lf, t.repl[0]}}];
        op.conj_a = P.vals[op.a].conj;
        op.are = t.coef_re; op.aim = t.coef_im;
        op.node = parent_of_leaf[lf];
        P.emit(P.fwd, op);
    }
    // internal nodes
    int lpos = 0;
    for (int j = 0; j < nn; ++j) {
        int ca = T->node_children[2 * j], cb = T->node_children[2 * j + 1];
        int self = nl + j;
        B.cur_node = j;
        if (ca < 0 || cb < 0 || ca >= self || cb >= self || ca == cb) return tne_fail(ctx, TNE_ERR_INVALID, "node %d has invalid children (%d, %d); nodes must be in post-order", j, ca, cb);
        NodeInfo& nd = B.nodes[self];
        NodeInfo& na = B.nodes[ca];
        NodeInfo& nb = B.nodes[cb];
        if (na.leaves.empty() || nb.leaves.empty()) return tne_fail(ctx, TNE_ERR_INVALID, "node %d uses a child twice", j);
        nd.leaves.resize(na.leaves.size() + nb.leaves.size());
        std::merge(na.leaves.begin(), na.leaves.end(), nb.leaves.begin(), nb.leaves.end(), nd.leaves.begin());
        int rk = T->node_out_rank[j];
        for (int q = 0; q < rk; ++q) {
            int32_t l = T->node_out_labels[lpos + q];
            if (find_in(B.sliced, l) < 0) nd.out_labels.push_back(l);
        }
        lpos += rk;
        const bool is_root = (j == nn - 1);
        const std::vector<int32_t>* forced = is_root ? &nd.out_labels : nullptr;
        // base
        nd.var[Key()] = B.contract_vals(na.var[Key()], nb.var[Key()], nd.out_labels, -1, 1.0, 0.0, forced);
        // partial patterns and completions
        std::map<std::pair<int, int>, std::pair<double, double>> closing;  // (val a, val b) -> summed coefficient
        for (int k = 0; k < B.n_terms; ++k) {
            const tne_term& t = B.terms[k];
            if (t.n_repl == 0) continue;
            Key r = B.restrict_key(k, nd.leaves);
            if (r.empty()) continue;
            Key ra = B.restrict_key(k, na.leaves), rb = B.restrict_key(k, nb.leaves);
            bool complete = (int)r.size() == t.n_repl;
            if (complete) {
                if ((int)ra.size() == t.n_repl || (int)rb.size() == t.n_repl) continue;  // already inside a child's sigma
                auto& c = closing[{na.var.at(ra), nb.var.at(rb)}];
                c.first += t.coef_re; c.second += t.coef_im;
            } else if (!nd.var.count(r)) {
                nd.var[r] = B.contract_vals(na.var.at(ra), nb.var.at(rb), nd.out_labels, -1, 1.0, 0.0, forced);
            }
        }
        auto sigma_add = [&](int va, int vb, double re, double im) {
            if (nd.sigma < 0) nd.sigma = B.contract_vals(va, vb, nd.out_labels, -1, re, im, forced);
            else B.contract_vals(va, vb, nd.out_labels, nd.sigma, re, im);
        };
        if (na.sigma >= 0) sigma_add(na.sigma, nb.var[Key()], 1.0, 0.0);
        if (nb.sigma >= 0) sigma_add(na.var[Key()], nb.sigma, 1.0, 0.0);
        for (auto& kv : closing) sigma_add(kv.first.first, kv.first.second, kv.second.first, kv.second.second);
        na.leaves.clear(); nb.leaves.clear();  // children consumed (a tree: each node has one parent)
        na.leaves.shrink_to_fit();
    }
    NodeInfo& root = B.nodes[nl + nn - 1];
    int result = root.var[Key()];
    if (B.n_terms > 0) {
        // terms without replacements scale the base network
        double cre = 0, cim = 0;
        bool any0 = false;
        for (int k = 0; k < B.n_terms; ++k) if (B.terms[k].n_repl == 0) { cre += B.terms[k].coef_re; cim += B.terms[k].coef_im; any0 = true; }
        if (root.sigma < 0) root.sigma = P.new_val(P.vals[result].labels, P.vals[result].dims);
        if (any0 || !P.written[root.sigma]) {
            Op op;
            op.kind = OP_AXPY;
            op.out = root.sigma; op.a = result;
            op.are = any0 ? cre : 0.0; op.aim = any0 ? cim : 0.0;
            op.node = nn - 1;
            P.emit(P.fwd, op);
        }
        result = root.sigma;
    }
    *root_out = result;
    (void)root_is_output;
    // dead-code elimination: drop forward ops whose results never reach the root
    {
        std::vector<bool> live(P.vals.size(), false);
        live[result] = true;
        std::vector<Op> kept;
        for (int i = (int)P.fwd.size() - 1; i >= 0; --i) {
            const Op& op = P.fwd[i];
            if (!live[op.out]) continue;
            if (op.a >= 0) live[op.a] = true;
            if (op.b >= 0) live[op.b] = true;
            kept.push_back(op);
        }
        std::reverse(kept.begin(), kept.end());
        // the first surviving writer of a value must overwrite, the rest accumulate
        std::vector<bool> seen(P.vals.size(), false);
        for (auto& op : kept) { op.acc = seen[op.out]; seen[op.out] = true; }
        P.fwd.swap(kept);
    }
    return TNE_OK;
}

// reverse-mode source transformation of the forward op list
void build_backward(Program& P, int root, int seed_val) {
    P.vals[root].grad = seed_val;
    // needs_grad propagation (forward order)
    for (auto& op : P.fwd) {
        bool ng = (op.a >= 0 && P.vals[op.a].needs_grad) || (op.b >= 0 && P.vals[op.b].needs_grad);
        if (ng) P.vals[op.out].needs_grad = true;
    }
    auto grad_of = [&](int v) {
        if (P.vals[v].grad < 0) {
            Val g;
            g.labels = P.vals[v].labels;
            g.dims = P.vals[v].dims;
            g.nelem = P.vals[v].nelem;
            P.vals.push_back(g);
            P.written.push_back(false);
            P.vals[v].grad = (int)P.vals.size() - 1;
        }
        return P.vals[v].grad;
    };
    for (int i = (int)P.fwd.size() - 1; i >= 0; --i) {
        const Op op = P.fwd[i];
        int dout = P.vals[op.out].grad;
        if (dout < 0 || !P.written[dout]) continue;  // no cotangent reaches this op
        if (op.kind == OP_AXPY) {
            if (!P.vals[op.a].needs_grad) continue;
            Op g;
            g.kind = OP_AXPY; g.backward = true; g.node = op.node; g.seg = op.seg;
            g.out = grad_of(op.a); g.a = dout;
            if (op.conj_a) { g.conj_a = true; g.are = op.are; g.aim = op.aim; }       // a' = conj(a): da = conj(conj(alpha) dout)
            else { g.conj_a = false; g.are = op.are; g.aim = -op.aim; }
            P.emit(P.bwd, g);
            continue;
        }
        // c (+)= alpha * fa(a) * fb(b)
        if (P.vals[op.a].needs_grad) {
            Op g;
            g.kind = OP_CONTRACT; g.backward = true; g.node = op.node; g.seg = op.seg;
            g.out = grad_of(op.a); g.a = dout; g.b = op.b;
            if (op.conj_a) { g.conj_a = true; g.conj_b = op.conj_b; g.are = op.are; g.aim = op.aim; }
            else { g.conj_a = false; g.conj_b = !op.conj_b; g.are = op.are; g.aim = -op.aim; }
            P.emit(P.bwd, g);
        }
        if (P.vals[op.b].needs_grad) {
            Op g;
            g.kind = OP_CONTRACT; g.backward = true; g.node = op.node; g.seg = op.seg;
            g.out = grad_of(op.b); g.a = op.a; g.b = dout;
            if (op.conj_b) { g.conj_b = true; g.conj_a = op.conj_a; g.are = op.are; g.aim = op.aim; }
            else { g.conj_b = false; g.conj_a = !op.conj_a; g.are = op.are; g.aim = -op.aim; }
            P.emit(P.bwd, g);
        }
    }
}


// ---------------------------------------------------------------------------------------
// phases: the executable form of a (segmented) program
// ---------------------------------------------------------------------------------------
struct PView {               // how one value looks inside one phase (segment-local slicing applied)
    std::vector<int32_t> labels;
    std::vector<int64_t> dims, strides;
    std::vector<int64_t> sl_stride;   // element stride per segment-sliced label (0: label absent)
    bool temp = false;                // phase-local storage (sliced labels removed), else checkpoint / external
    int64_t off = 0;                  // byte offset: temp arena (temp) or checkpoint arena
    bool dep = false;                 // differs from slice to slice
    int64_t nelem = 1;
    bool tape = false;                // kept from the forward phase for the reverse phase of a retained (segment, slice) unit
    int64_t tape_off = 0;             // byte offset inside the unit's tape block
};

struct PhaseOp {
    Op op;
    ContractPlan plan;
    bool first_slice_only = false;
    bool skip_if_retained = false;     // recompute op whose result a retained unit still has from its forward phase
    bool dead = false;                 // merged into a later op
    // further contributions into the same output, fused into one launch (gett_absorb_multi)
    struct Extra { int a, b; ContractPlan plan; };
    std::vector<Extra> extra;
    // fused site absorption C = (A . B1) . B2 (tne_site2.cu): this op is the second step, s2 holds both
    std::shared_ptr<Site2Desc> s2;
    int s2_a = -1, s2_b1 = -1, s2_b2 = -1;
    double s2_flops = 0, s2_bytes = 0;
    // fused reverse pass of a site (tne_site2b.cu): this op is dA = dT . B1^H; the producer of dT and the leaf-gradient op
    // dB1 += A^H . dT are folded in (op.out = dA)
    std::shared_ptr<Site2BwdDesc> s2b;
    int s2b_dC = -1, s2b_B2 = -1, s2b_B1 = -1, s2b_A = -1, s2b_dB1 = -1;
    bool s2b_add = false, s2b_final = true;    // CTA partials of the leaf gradient: add to the workspace / reduce into dB1 afterwards
    std::vector<int> inputs() const {
        if (s2) return {s2_a, s2_b1, s2_b2};
        if (s2b) return {s2b_dC, s2b_B2, s2b_B1, s2b_A};
        std::vector<int> v = {op.a, op.b};
        for (auto& ex : extra) { v.push_back(ex.a); v.push_back(ex.b); }
        return v;
    }
};

struct Phase {
    int seg = 0;
    bool backward = false;
    std::vector<int32_t> sliced;
    std::vector<int64_t> ssz;
    int64_t nsl = 1;
    std::vector<PhaseOp> ops;
    std::map<int, PView> views;
    std::vector<int> zero_at_start;   // checkpoint values this phase accumulates into
    // how the checkpoints this phase produced are combined over the ranks when its slices are dealt to them (seg_shard):
    // 0 all-reduce; 1 reduce-scatter (every consumer reads only the slices its rank owns); 2 all-gather (every rank wrote
    // only the slices it owns).  block / nblocks: the contiguous slice blocks, block b owned by rank b % world.
    struct Comm { int kind = 0; int64_t block = 0, nblocks = 0; };
    std::map<int, Comm> comm;
    int64_t temp_bytes = 0;
    // tape retention: byte offset of the tape block of slice `it` of this segment (-1: the unit is recomputed), shared by
    // the forward and the reverse phase of the segment; tape_unit_bytes: size of one block
    std::vector<int64_t> unit_tape_base;
    int64_t tape_unit_bytes = 0;
};

struct Interval { int id; int64_t bytes; int first, last; int64_t off; };

// greedy-by-size offset assignment over exact lifetimes; returns the peak
int64_t plan_intervals(std::vector<Interval>& iv) {
    std::vector<int> order(iv.size());
    for (size_t i = 0; i < iv.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int x, int y) {
        if (iv[x].bytes != iv[y].bytes) return iv[x].bytes > iv[y].bytes;
        return iv[x].first < iv[y].first;
    });
    std::vector<int> placed;
    int64_t peak = 0;
    for (int v : order) {
        std::vector<std::pair<int64_t, int64_t>> busy;
        for (int u : placed)
            if (!(iv[u].last < iv[v].first || iv[v].last < iv[u].first)) busy.emplace_back(iv[u].off, iv[u].off + iv[u].bytes);
        std::sort(busy.begin(), busy.end());
        int64_t off = 0;
        for (auto& b : busy) {
            if (b.first - off >= iv[v].bytes) break;
            off = std::max(off, b.second);
        }
        iv[v].off = off;
        placed.push_back(v);
        peak = std::max(peak, off + iv[v].bytes);
    }
    return peak;
}

inline int64_t round512(int64_t b) { return (b + 511) / 512 * 512; }

struct Exec {
    Program& P;
    tne_ctx* ctx;
    std::vector<Phase> phases;
    int64_t persist_bytes = 0, temp_bytes = 0, tape_bytes = 0;
    int tape_units = 0, tape_units_possible = 0;
    std::map<int32_t, int64_t> label_size;
    std::set<int> ever_zeroed;   // checkpoints are cleared by the first phase that writes them only
    Exec(Program& p) : P(p), ctx(p.ctx) {}

    // Layout of workspace tensors: column-major; with option "pad" > 0 a leg whose stride would reach
    // 64 KiB gets that many extra elements of stride (an experiment against HBM channel aliasing of the
    // huge power-of-two strides the sweep produces; measured to bring nothing on B200, so off by default).
    // Returns the extent in elements.
    int64_t padded_strides(const std::vector<int64_t>& dims, std::vector<int64_t>& st) const {
        st.resize(dims.size());
        const int64_t pad = ctx->opt_pad;
        int64_t acc = 1;
        for (size_t i = 0; i < dims.size(); ++i) {
            if (pad > 0 && i > 0 && dims[i] > 1 && acc * (int64_t)sizeof(cplx) >= (64 << 10)) acc += pad;
            st[i] = acc;
            acc *= dims[i];
        }
        return acc;
    }
    int64_t extent_of(const Val& v) const {
        std::vector<int64_t> st;
        return padded_strides(v.dims, st);
    }
    void full_strides(const Val& v, std::vector<int64_t>& st) const {
        if (!v.strides.empty()) { st = v.strides; return; }
        if (v.ext) {
            st.resize(v.dims.size());
            int64_t acc = 1;
            for (size_t i = 0; i < v.dims.size(); ++i) { st[i] = acc; acc *= v.dims[i]; }
            return;
        }
        padded_strides(v.dims, st);
    }

    // segment bookkeeping: op.seg from the node -> segment map; persistence of values
    int assign_segments(const tne_tree* T) {
        const int nn = T->n_nodes;
        int S = std::max(1, (int)T->n_segments);
        std::vector<int> seg_of_node(nn, 0);
        if (T->n_segments > 1) {
            if (!T->seg_first_node || T->seg_first_node[0] != 0) return tne_fail(ctx, TNE_ERR_INVALID, "seg_first_node[0] must be 0");
            for (int s = 0; s < S; ++s) {
                int lo = T->seg_first_node[s], hi = s + 1 < S ? T->seg_first_node[s + 1] : nn;
                if (lo < 0 || hi > nn || lo >= hi) return tne_fail(ctx, TNE_ERR_INVALID, "segment %d covers no nodes", s);
                for (int j = lo; j < hi; ++j) seg_of_node[j] = s;
            }
        }
        for (auto& op : P.fwd) op.seg = seg_of_node[op.node];
        std::stable_sort(P.fwd.begin(), P.fwd.end(), [](const Op& x, const Op& y) { return x.seg < y.seg; });
        return TNE_OK;
    }

    void mark_persistent() {
        std::vector<int> seg_seen(P.vals.size(), -1);
        auto touch = [&](int v, int seg) {
            if (v < 0 || P.vals[v].ext) return;
            if (seg_seen[v] < 0) seg_seen[v] = seg;
            else if (seg_seen[v] != seg) P.vals[v].persistent = true;
        };
        for (auto& op : P.fwd) { touch(op.out, op.seg); touch(op.a, op.seg); touch(op.b, op.seg); }
        for (auto& op : P.bwd) { touch(op.out, op.seg); touch(op.a, op.seg); touch(op.b, op.seg); }
    }

    int make_view(Phase& ph, int v) {
        if (v < 0 || ph.views.count(v)) return TNE_OK;
        const Val& val = P.vals[v];
        PView pv;
        pv.temp = !val.ext && !val.persistent;
        std::vector<int64_t> st;
        full_strides(val, st);
        pv.sl_stride.assign(ph.sliced.size(), 0);
        for (size_t q = 0; q < val.labels.size(); ++q) {
            int s = find_in(ph.sliced, val.labels[q]);
            if (s >= 0) {
                pv.dep = true;
                if (!pv.temp) pv.sl_stride[s] = st[q];
                continue;
            }
            pv.labels.push_back(val.labels[q]);
            pv.dims.push_back(val.dims[q]);
            pv.strides.push_back(st[q]);
        }
        for (auto d : pv.dims) pv.nelem *= d;
        if (pv.temp) pv.nelem = padded_strides(pv.dims, pv.strides);  // dense (padded) in its reduced shape; nelem = extent
        ph.views.emplace(v, std::move(pv));
        return TNE_OK;
    }

    TensorRef ref_of(const Phase& ph, int v, bool conj) const {
        const PView& pv = ph.views.at(v);
        TensorRef t;
        t.labels = pv.labels; t.dims = pv.dims; t.strides = pv.strides; t.conj = conj;
        return t;
    }

    int build_phase(Phase& ph, const std::vector<Op>& ops) {
        for (auto& op : ops) { TNE_TRY(make_view(ph, op.out)); TNE_TRY(make_view(ph, op.a)); TNE_TRY(make_view(ph, op.b)); }
        std::set<int> seen;
        for (auto op : ops) {
            PView& vo = ph.views.at(op.out);
            bool in_dep = (op.a >= 0 && ph.views.at(op.a).dep) || (op.b >= 0 && ph.views.at(op.b).dep);
            bool out_view_dep = false;
            for (auto s : vo.sl_stride) if (s != 0) out_view_dep = true;
            PhaseOp po;
            po.first_slice_only = false;
            if (vo.temp) {
                op.acc = seen.count(op.out) > 0;
                seen.insert(op.out);
                if (P.vals[op.out].labels.size() != vo.labels.size()) out_view_dep = true;
            } else {
                op.acc = true;  // checkpoints are cleared at phase start, external outputs at program start
                if (!P.vals[op.out].ext && !ever_zeroed.count(op.out)) { ever_zeroed.insert(op.out); ph.zero_at_start.push_back(op.out); }
                if (!in_dep && !out_view_dep && ph.nsl > 1) po.first_slice_only = true;
            }
            if (in_dep || out_view_dep) vo.dep = true;
            po.op = op;
            TensorRef C = ref_of(ph, op.out, false);
            TensorRef A = ref_of(ph, op.a, op.conj_a);
            TensorRef Bt;
            if (op.kind == OP_CONTRACT) Bt = ref_of(ph, op.b, op.conj_b);
            TNE_TRY(contract_plan(ctx, A, Bt, C, op.acc, op.are, op.aim, &po.plan));
            if (ctx->opt_debug > 2 && !po.plan.empty) {
                const GettDesc& d = po.plan.d;
                fprintf(stderr, "[tne-op] seg %d %s%s kind %d kernel %d acc %d swap %d M %lld N %lld K %lld B %lld conj %d%d |", ph.seg,
                        op.backward ? "bwd" : "fwd", op.recompute ? "(re)" : "", op.kind, po.plan.kernel, (int)op.acc, (int)po.plan.swapped,
                        d.Msz, d.Nsz, d.Ksz, d.Bsz, d.conjA, d.conjB);
                auto dump = [&](const char* nm, const LegSet& L) {
                    fprintf(stderr, " %s:", nm);
                    for (int i = 0; i < L.n; ++i) fprintf(stderr, " %d(%lld,%lld,%lld)", L.size[i], L.s0[i], L.s1[i], L.s2[i]);
                };
                dump("M[a,c]", d.M); dump("N[b,c]", d.N); dump("K[a,b]", d.K); dump("Bt[a,b,c]", d.Bt);
                fprintf(stderr, "\n");
            }
            ph.ops.push_back(std::move(po));
        }
        // fuse the two steps of a site, T = A . B1 then C = T . B2, when T has no other reader in this phase
        if (!ctx->opt_no_site2) {
            std::map<int, int> reads, producer;
            for (int i = 0; i < (int)ph.ops.size(); ++i) {
                const Op& op = ph.ops[i].op;
                if (op.a >= 0) reads[op.a]++;
                if (op.b >= 0) reads[op.b]++;
                producer[op.out] = producer.count(op.out) ? -2 : i;   // -2: several writers
            }
            for (int j = 0; j < (int)ph.ops.size(); ++j) {
                PhaseOp& pj = ph.ops[j];
                if (pj.op.kind != OP_CONTRACT || pj.first_slice_only || pj.plan.empty) continue;
                for (int side = 0; side < 2; ++side) {
                    const int T = side == 0 ? pj.op.a : pj.op.b, b2 = side == 0 ? pj.op.b : pj.op.a;
                    const bool conjT = side == 0 ? pj.op.conj_a : pj.op.conj_b, conjB2 = side == 0 ? pj.op.conj_b : pj.op.conj_a;
                    if (T < 0 || b2 < 0 || conjT) continue;
                    auto itp = producer.find(T);
                    if (itp == producer.end() || itp->second < 0 || itp->second >= j || reads[T] != 1) continue;
                    PhaseOp& pi = ph.ops[itp->second];
                    if (pi.dead || pi.op.kind != OP_CONTRACT || pi.op.acc || pi.first_slice_only || pi.plan.empty || pi.s2) continue;
                    if (!ph.views.at(T).temp) continue;
                    // the big operand of the first step is A
                    const bool a_big = ph.views.at(pi.op.a).nelem >= ph.views.at(pi.op.b).nelem;
                    const int vA = a_big ? pi.op.a : pi.op.b, vB1 = a_big ? pi.op.b : pi.op.a;
                    TensorRef rA = ref_of(ph, vA, a_big ? pi.op.conj_a : pi.op.conj_b);
                    TensorRef rB1 = ref_of(ph, vB1, a_big ? pi.op.conj_b : pi.op.conj_a);
                    TensorRef rB2 = ref_of(ph, b2, conjB2);
                    TensorRef rC = ref_of(ph, pj.op.out, false);
                    Site2Desc sd;
                    if (!site2_plan(ctx, rA, rB1, rB2, rC, pj.op.acc, pi.op.are, pi.op.aim, pj.op.are, pj.op.aim, &sd)) continue;
                    pj.s2 = std::make_shared<Site2Desc>(sd);
                    pj.s2_a = vA; pj.s2_b1 = vB1; pj.s2_b2 = b2;
                    pj.s2_flops = pi.plan.flops + pj.plan.flops;
                    pj.s2_bytes = 16.0 * ((double)ph.views.at(vA).nelem + (double)ph.views.at(pj.op.out).nelem);
                    pi.dead = true;
                    break;
                }
            }
            std::vector<PhaseOp> kept;
            for (auto& po : ph.ops) if (!po.dead) kept.push_back(std::move(po));
            ph.ops.swap(kept);
        }
        // fuse the three reverse ops of a site -- dT = dC . B2^H (single writer), dA = dT . B1^H, dB1 += A^H . dT -- when dT has
        // no other reader: dT then never exists in memory (tne_site2b.cu)
        if (!ctx->opt_no_site2b) {
            std::map<int, std::vector<int>> writers, readers;
            for (int i = 0; i < (int)ph.ops.size(); ++i) {
                for (int v : ph.ops[i].inputs()) if (v >= 0) readers[v].push_back(i);
                writers[ph.ops[i].op.out].push_back(i);
            }
            auto plain_bwd = [&](const PhaseOp& po) {
                return po.op.kind == OP_CONTRACT && po.op.backward && !po.s2 && !po.s2b && !po.dead && !po.first_slice_only && !po.plan.empty && po.extra.empty();
            };
            for (int j = 0; j < (int)ph.ops.size(); ++j) {
                PhaseOp& pj = ph.ops[j];
                if (!plain_bwd(pj)) continue;
                for (int side = 0; side < 2 && !pj.s2b; ++side) {
                    const int dT = side == 0 ? pj.op.a : pj.op.b, vB1 = side == 0 ? pj.op.b : pj.op.a;
                    const bool cT2 = side == 0 ? pj.op.conj_a : pj.op.conj_b, cB1 = side == 0 ? pj.op.conj_b : pj.op.conj_a;
                    if (dT < 0 || vB1 < 0 || !ph.views.at(dT).temp) continue;
                    auto wi = writers.find(dT);
                    auto ri = readers.find(dT);
                    if (wi == writers.end() || wi->second.size() != 1 || ri == readers.end() || ri->second.size() != 2) continue;
                    const int i = wi->second[0];
                    const int k = ri->second[0] == j ? ri->second[1] : ri->second[0];
                    if (i >= j || k == j || k <= i) continue;
                    PhaseOp& pi = ph.ops[i];
                    PhaseOp& pk = ph.ops[k];
                    if (!plain_bwd(pi) || !plain_bwd(pk) || pi.op.acc) continue;
                    // producer of dT: the big operand is the cotangent dC, the small one the ket tensor
                    const bool a_big = ph.views.at(pi.op.a).nelem >= ph.views.at(pi.op.b).nelem;
                    const int vdC = a_big ? pi.op.a : pi.op.b, vB2 = a_big ? pi.op.b : pi.op.a;
                    const bool cC = a_big ? pi.op.conj_a : pi.op.conj_b, cB2 = a_big ? pi.op.conj_b : pi.op.conj_a;
                    // leaf gradient: the operand that is not dT is the forward boundary A; the output must not be phase-local
                    if (pk.op.a != dT && pk.op.b != dT) continue;
                    const int vA = pk.op.a == dT ? pk.op.b : pk.op.a;
                    const bool cA = pk.op.a == dT ? pk.op.conj_b : pk.op.conj_a, cT3 = pk.op.a == dT ? pk.op.conj_a : pk.op.conj_b;
                    if (vA < 0 || vA == dT || ph.views.at(pk.op.out).temp) continue;
                    TensorRef rdC = ref_of(ph, vdC, cC), rB2 = ref_of(ph, vB2, cB2), rB1 = ref_of(ph, vB1, cB1), rA = ref_of(ph, vA, cA);
                    TensorRef rdA = ref_of(ph, pj.op.out, false), rdB1 = ref_of(ph, pk.op.out, false);
                    Site2BwdDesc sd;
                    if (!site2b_plan(ctx, rdC, rB2, rB1, rA, rdA, rdB1, &sd)) continue;
                    sd.conjC = cC; sd.conjB2 = cB2; sd.conjT2 = cT2; sd.conjB1 = cB1; sd.conjA = cA; sd.conjT3 = cT3;
                    sd.accumulate = pj.op.acc;
                    sd.a1re = pi.op.are; sd.a1im = pi.op.aim; sd.a2re = pj.op.are; sd.a2im = pj.op.aim; sd.a3re = pk.op.are; sd.a3im = pk.op.aim;
                    pj.s2b = std::make_shared<Site2BwdDesc>(sd);
                    pj.s2b_dC = vdC; pj.s2b_B2 = vB2; pj.s2b_B1 = vB1; pj.s2b_A = vA; pj.s2b_dB1 = pk.op.out;
                    pj.s2_flops = pi.plan.flops + pj.plan.flops + pk.plan.flops;
                    pj.s2_bytes = 16.0 * ((double)ph.views.at(vdC).nelem + (double)ph.views.at(vA).nelem + (double)ph.views.at(pj.op.out).nelem);
                    pi.dead = true; pk.dead = true;
                }
            }
            std::vector<PhaseOp> kept;
            for (auto& po : ph.ops) if (!po.dead) kept.push_back(std::move(po));
            ph.ops.swap(kept);
            // consecutive fused launches into one leaf gradient (the variants of a site) share the CTA-partial workspace: the
            // first overwrites it, the others add, one reduction into dB1 after the last
            for (size_t q = 0; q < ph.ops.size(); ++q) {
                if (!ph.ops[q].s2b) continue;
                size_t e = q;
                while (e + 1 < ph.ops.size() && ph.ops[e + 1].s2b && ph.ops[e + 1].s2b_dB1 == ph.ops[q].s2b_dB1) ++e;
                for (size_t z = q; z <= e; ++z) { ph.ops[z].s2b_add = z > q; ph.ops[z].s2b_final = z == e; }
                q = e;
            }
        }
        // fuse consecutive writers of one output that run the absorb kernel with identical geometry: the group
        // moves to the position of its last member (all inputs are ready there), C is written once
        if (!ctx->opt_no_fuse) {
            std::vector<PhaseOp> merged;
            std::map<int, int> open;   // output value -> index in `merged` of a group that may still grow
            for (auto& po : ph.ops) {
                const int out = po.op.out;
                for (int v : po.inputs()) if (v >= 0) open.erase(v);   // a reader closes the group of the value it reads
                auto it = open.find(out);
                if (po.s2 || po.s2b) { merged.push_back(po); open.erase(out); if (po.s2b) open.erase(po.s2b_dB1); continue; }
                if (it != open.end() && po.op.kind == OP_CONTRACT && !po.plan.empty) {
                    PhaseOp& g = merged[it->second];
                    if ((int)g.extra.size() + 1 < TNE_MULTI_MAX && g.first_slice_only == po.first_slice_only &&
                        contract_multi_compatible(g.plan, po.plan)) {
                        PhaseOp moved = g;
                        g.dead = true;
                        moved.extra.push_back({po.op.a, po.op.b, po.plan});
                        merged.push_back(std::move(moved));
                        open[out] = (int)merged.size() - 1;
                        continue;
                    }
                }
                merged.push_back(po);
                if (po.op.kind == OP_CONTRACT && !po.plan.empty && contract_multi_compatible(po.plan, po.plan)) open[out] = (int)merged.size() - 1;
                else open.erase(out);
            }
            std::vector<PhaseOp> kept;
            for (auto& po : merged) if (!po.dead) kept.push_back(std::move(po));
            ph.ops.swap(kept);
        }
        // temp arena of this phase
        std::vector<Interval> iv;
        std::map<int, int> slot;
        for (int i = 0; i < (int)ph.ops.size(); ++i) {
            const Op& op = ph.ops[i].op;
            std::vector<int> ids = {op.out, op.a, op.b};
            for (auto& ex : ph.ops[i].extra) { ids.push_back(ex.a); ids.push_back(ex.b); }
            if (ph.ops[i].s2) { ids = {op.out, ph.ops[i].s2_a, ph.ops[i].s2_b1, ph.ops[i].s2_b2}; }
            if (ph.ops[i].s2b) { ids = {op.out, ph.ops[i].s2b_dC, ph.ops[i].s2b_B2, ph.ops[i].s2b_B1, ph.ops[i].s2b_A, ph.ops[i].s2b_dB1}; }
            for (int v : ids) {
                if (v < 0 || !ph.views.at(v).temp) continue;
                auto it = slot.find(v);
                if (it == slot.end()) { slot[v] = (int)iv.size(); iv.push_back({v, round512(ph.views.at(v).nelem * (int64_t)sizeof(cplx)), i, i, 0}); }
                else iv[it->second].last = i;
            }
        }
        ph.temp_bytes = plan_intervals(iv);
        for (auto& x : iv) ph.views.at(x.id).off = x.off;
        return TNE_OK;
    }

    int build(const tne_tree* T, bool has_grad) {
        const int S = std::max(1, (int)T->n_segments);
        // segment-local sliced labels
        std::vector<std::vector<int32_t>> seg_sliced(S);
        if (T->n_segments > 1 && T->seg_sliced_off)
            for (int s = 0; s < S; ++s)
                for (int q = T->seg_sliced_off[s]; q < T->seg_sliced_off[s + 1]; ++q) seg_sliced[s].push_back(T->seg_sliced_labels[q]);
        for (auto& v : P.vals) for (size_t q = 0; q < v.labels.size(); ++q) label_size[v.labels[q]] = v.dims[q];
        mark_persistent();
        std::vector<std::vector<Op>> fseg(S), bseg(S);
        for (auto& op : P.fwd) fseg[op.seg].push_back(op);
        for (auto& op : P.bwd) bseg[op.seg].push_back(op);
        auto new_phase = [&](int s, bool backward) {
            Phase ph;
            ph.seg = s; ph.backward = backward;
            for (auto l : seg_sliced[s]) {
                auto it = label_size.find(l);
                if (it == label_size.end()) continue;  // label does not occur (e.g. sliced globally)
                ph.sliced.push_back(l); ph.ssz.push_back(it->second); ph.nsl *= it->second;
            }
            return ph;
        };
        // forward phases; the last segment keeps its tape and runs its reverse ops in the same phase
        for (int s = 0; s < S; ++s) {
            std::vector<Op> ops = fseg[s];
            bool fused_bwd = has_grad && s == S - 1;
            if (fused_bwd) ops.insert(ops.end(), bseg[s].begin(), bseg[s].end());
            if (ops.empty()) continue;
            Phase ph = new_phase(s, fused_bwd);
            TNE_TRY(build_phase(ph, ops));
            phases.push_back(std::move(ph));
        }
        // reverse phases with recomputation of the segment's own forward values
        for (int s = S - 2; s >= 0 && has_grad; --s) {
            if (bseg[s].empty()) continue;
            std::set<int> need;
            for (auto& op : bseg[s]) {
                int ins[2] = {op.a, op.b};
                for (int v : ins) if (v >= 0 && !P.vals[v].ext && !P.vals[v].persistent) need.insert(v);
            }
            // only values produced by forward ops of this segment can be recomputed
            std::set<int> fwd_outs;
            for (auto& op : fseg[s]) fwd_outs.insert(op.out);
            std::vector<Op> re;
            for (int i = (int)fseg[s].size() - 1; i >= 0; --i) {
                const Op& op = fseg[s][i];
                if (!need.count(op.out) || P.vals[op.out].persistent || P.vals[op.out].ext) continue;
                Op r = op;
                r.recompute = true;
                re.push_back(r);
                int ins[2] = {op.a, op.b};
                for (int v : ins) if (v >= 0 && fwd_outs.count(v)) need.insert(v);
            }
            std::reverse(re.begin(), re.end());
            Phase ph = new_phase(s, true);
            if (ph.nsl > 1)
                for (auto& op : re) {
                    int ins[2] = {op.a, op.b};
                    for (int v : ins)
                        if (v >= 0 && P.vals[v].persistent && fwd_outs.count(v))
                            return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "segment %d re-reads its own checkpoint inside a sliced segment", s);
                }
            re.insert(re.end(), bseg[s].begin(), bseg[s].end());
            TNE_TRY(build_phase(ph, re));
            phases.push_back(std::move(ph));
        }
        // checkpoint arena over the phase timeline
        std::vector<Interval> iv;
        std::map<int, int> slot;
        for (int pi = 0; pi < (int)phases.size(); ++pi)
            for (auto& kv : phases[pi].views) {
                int v = kv.first;
                if (!P.vals[v].persistent) continue;
                auto it = slot.find(v);
                if (it == slot.end()) { slot[v] = (int)iv.size(); iv.push_back({v, round512(extent_of(P.vals[v]) * (int64_t)sizeof(cplx)), pi, pi, 0}); }
                else iv[it->second].last = pi;
            }
        persist_bytes = plan_intervals(iv);
        for (auto& x : iv) P.vals[x.id].offset = x.off;
        for (auto& ph : phases) temp_bytes = std::max(temp_bytes, ph.temp_bytes);
        return TNE_OK;
    }

    // The slices of value v, as phase ph loops over them, are contiguous blocks in iteration order.  The sliced labels of
    // the phase that v carries must be the FIRST ones of the phase's list (the fastest digits of the iteration counter, so
    // that block j is touched by the iterations it = j mod nblocks only); `whole`: v must carry all of them.
    bool contiguous_slices(const Phase& ph, int v, int64_t* block, int64_t* nblocks, bool whole) const {
        auto it = ph.views.find(v);
        if (it == ph.views.end() || ph.nsl <= 1) return false;
        const PView& pv = it->second;
        size_t q = 0;
        int64_t nb = 1;
        while (q < ph.ssz.size() && pv.sl_stride[q] != 0) { nb *= ph.ssz[q]; ++q; }
        for (size_t z = q; z < ph.ssz.size(); ++z) if (pv.sl_stride[z] != 0) return false;
        if (q == 0 || (whole && q != ph.ssz.size())) return false;
        const int64_t ext = extent_of(P.vals[v]);
        if (ext % nb != 0) return false;
        int64_t want = ext / nb;
        *block = want;
        *nblocks = nb;
        for (size_t z = 0; z < q; ++z) {
            if (pv.sl_stride[z] != want) return false;
            want *= ph.ssz[z];
        }
        return true;
    }

    // Decide, per checkpoint a sliced phase produces, which collective combines it over the ranks (option rs_comm).
    // A forward checkpoint is a sum over the producer's slices of a tensor that carries the NEXT column's sliced legs: if
    // every later phase that touches it loops over those legs (and they are its slowest ones), a rank only ever reads the
    // slices it owns -> reduce-scatter.  A cotangent checkpoint carries the producer's own sliced legs: every rank writes
    // only the slices it owns -> all-gather.  Anything else keeps the all-reduce.  Half the NVLink traffic either way.
    void plan_comm(int world) {
        for (size_t pi = 0; pi < phases.size(); ++pi) {
            Phase& ph = phases[pi];
            ph.comm.clear();
            if (!ctx->opt_rs_comm || world <= 1 || ph.nsl <= 1) continue;
            for (int v : ph.zero_at_start) {
                Phase::Comm c;
                const PView& pv = ph.views.at(v);
                bool all_sliced = !pv.sl_stride.empty();
                for (auto st : pv.sl_stride) if (st == 0) all_sliced = false;
                int64_t block = 0, nb = 0;
                if (all_sliced && ph.nsl % world == 0 && contiguous_slices(ph, v, &block, &nb, true)) {
                    c.kind = 2; c.block = block; c.nblocks = ph.nsl;
                }
                if (c.kind == 0) {
                    // every rank holds a partial sum of the WHOLE tensor (whatever it did not write is zero): the ranks only
                    // need the blocks they will read
                    bool ok = true, any = false;
                    for (size_t pj = 0; pj < phases.size() && ok; ++pj) {
                        if (pj == pi || !phases[pj].views.count(v)) continue;
                        const Phase& cons = phases[pj];
                        int64_t b2 = 0, n2 = 0;
                        if (!contiguous_slices(cons, v, &b2, &n2, false) || n2 % world != 0) { ok = false; break; }
                        for (int z : cons.zero_at_start) if (z == v) ok = false;          // a second writer
                        for (auto& po : cons.ops) if (po.op.out == v) ok = false;
                        if (any && (b2 != block || n2 != nb)) ok = false;
                        block = b2; nb = n2; any = true;
                    }
                    if (ok && any) { c.kind = 1; c.block = block; c.nblocks = nb; }
                }
                ph.comm[v] = c;
            }
        }
    }

    int site2_launch(const PhaseOp& po, const cplx* A, const cplx* B1, const cplx* B2, cplx* C) {
        if (!ctx->opt_profile) return site2_run(ctx, *po.s2, A, B1, B2, C);
        tne_ctx::ProfRec r;
        cudaEvent_t* ev[2] = {&r.e0, &r.e1};
        for (auto e : ev) {
            if (!ctx->event_pool.empty()) { *e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
            else TNE_CUDA(ctx, cudaEventCreate(e));
        }
        r.cls = 8 * 5; r.flops = po.s2_flops; r.bytes = po.s2_bytes;
        TNE_CUDA(ctx, cudaEventRecord(r.e0, ctx->stream));
        int rc = site2_run(ctx, *po.s2, A, B1, B2, C);
        TNE_CUDA(ctx, cudaEventRecord(r.e1, ctx->stream));
        ctx->prof.push_back(r);
        return rc;
    }

    int site2b_launch(const PhaseOp& po, const cplx* dC, const cplx* B2, const cplx* B1, const cplx* A, cplx* dA, cplx* dB1, bool skip_leaf) {
        const bool fin = po.s2b_final && !skip_leaf;
        if (!ctx->s2b_part && !ctx->plan_only)
            TNE_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->s2b_part), (size_t)site2b_part_elems(ctx) * sizeof(cplx)));
        if (!ctx->opt_profile) return site2b_run(ctx, *po.s2b, dC, B2, B1, A, dA, dB1, ctx->s2b_part, po.s2b_add, fin);
        tne_ctx::ProfRec r;
        cudaEvent_t* ev[2] = {&r.e0, &r.e1};
        for (auto e : ev) {
            if (!ctx->event_pool.empty()) { *e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
            else TNE_CUDA(ctx, cudaEventCreate(e));
        }
        r.cls = 8 * 6; r.flops = po.s2_flops; r.bytes = po.s2_bytes;
        TNE_CUDA(ctx, cudaEventRecord(r.e0, ctx->stream));
        int rc = site2b_run(ctx, *po.s2b, dC, B2, B1, A, dA, dB1, ctx->s2b_part, po.s2b_add, fin);
        TNE_CUDA(ctx, cudaEventRecord(r.e1, ctx->stream));
        ctx->prof.push_back(r);
        return rc;
    }

    cplx* base_ptr(const Phase& ph, int v, int64_t tape_base = -1) const {
        const Val& val = P.vals[v];
        if (val.ext) return val.ext + val.slice_base;
        const PView& pv = ph.views.at(v);
        char* ws = reinterpret_cast<char*>(ctx->ws);
        if (pv.tape && tape_base >= 0) return reinterpret_cast<cplx*>(ws + persist_bytes + temp_bytes + tape_base + pv.tape_off);
        return reinterpret_cast<cplx*>(pv.temp ? ws + persist_bytes + pv.off : ws + val.offset);
    }

    // Tape retention.  The reverse phase of a sliced segment re-runs the segment's forward ops for every slice because the
    // tapes of all (segment, slice) units do not fit together (860 GB at 6x6 D=4).  Whatever DOES fit in the memory that is
    // left keeps its tape: the forward values the reverse ops read (the boundaries before every site, not the twice-larger
    // bra intermediates) are written to a block of their own per retained unit and the unit's recompute ops are skipped.
    // One B200 holds ~10 of the 64 units of the benchmark; with the slices dealt to 8 ranks every rank holds ALL of its 8-9
    // units and the recompute (27 % of the executed flops) disappears.
    void plan_tape(int64_t budget, int world, int rank) {
        tape_bytes = 0; tape_units = 0; tape_units_possible = 0;
        if (budget <= 0) return;
        for (size_t pb = 0; pb < phases.size(); ++pb) {
            Phase& B = phases[pb];
            if (!B.backward) continue;
            bool has_re = false;
            for (auto& po : B.ops) if (po.op.recompute) has_re = true;
            if (!has_re) continue;
            size_t pf = phases.size();
            for (size_t q = 0; q < pb; ++q) if (phases[q].seg == B.seg && !phases[q].backward) pf = q;
            if (pf == phases.size()) continue;
            Phase& F = phases[pf];
            if (F.nsl != B.nsl) continue;
            // values the reverse ops (not the recompute ops) read that the forward phase materialises per slice
            std::set<int> fwd_written;
            for (auto& po : F.ops) fwd_written.insert(po.op.out);
            std::set<int> need;
            bool ok = true;
            for (auto& po : B.ops) {
                if (po.op.recompute) continue;
                for (int v : po.inputs()) {
                    if (v < 0) continue;
                    const PView& pv = B.views.at(v);
                    if (!pv.temp || (!pv.dep && B.nsl > 1)) continue;    // slice-independent values of a sliced segment are cheap: recomputed
                    bool made_here = false;        // produced by a recompute op of this phase (i.e. a forward value)?
                    for (auto& pr : B.ops) if (pr.op.recompute && pr.op.out == v) made_here = true;
                    if (!made_here) continue;      // a cotangent or another reverse value
                    if (!fwd_written.count(v) || !F.views.count(v) || !F.views.at(v).temp) { ok = false; break; }
                    need.insert(v);
                }
                if (!ok) break;
            }
            if (!ok || need.empty()) continue;
            int64_t off = 0;
            for (int v : need) {
                PView& vb = B.views.at(v);
                PView& vf = F.views.at(v);
                vb.tape = vf.tape = true;
                vb.tape_off = vf.tape_off = off;
                off += round512(vb.nelem * (int64_t)sizeof(cplx));
            }
            for (auto& po : B.ops)
                if (po.op.recompute && (B.views.at(po.op.out).dep || B.nsl == 1) && B.views.at(po.op.out).temp) po.skip_if_retained = true;
            F.tape_unit_bytes = B.tape_unit_bytes = off;
            F.unit_tape_base.assign(F.nsl, -1);
            const bool sharded = world > 1 && F.nsl > 1;
            for (int64_t it = 0; it < F.nsl; ++it) {
                if (sharded && it % world != rank) continue;
                ++tape_units_possible;
                if (tape_bytes + off > budget) continue;
                F.unit_tape_base[it] = tape_bytes;
                tape_bytes += off;
                ++tape_units;
            }
            B.unit_tape_base = F.unit_tape_base;
        }
    }

    // one line per launch group of the final (fused) phases: what runs, how often, with which algorithmic cost
    void dump_final() const {
        double tot_flops = 0;
        long long tot_launches = 0;
        for (auto& ph : phases)
            for (auto& po : ph.ops) {
                long long n = po.first_slice_only ? 1 : ph.nsl;
                if (po.skip_if_retained)
                    for (auto tb : ph.unit_tape_base) if (tb >= 0) --n;     // units that kept their tape do not recompute
                double fl = (po.s2 || po.s2b) ? po.s2_flops : po.plan.flops;
                for (auto& ex : po.extra) fl += ex.plan.flops;
                tot_flops += fl * (double)n;
                tot_launches += n;
            }
        fprintf(stderr, "[tne-plan] %zu phases, %lld launches, %.4e flops, workspace %.3f GiB (checkpoints %.3f + segment %.3f)\n", phases.size(),
                tot_launches, tot_flops, (persist_bytes + temp_bytes) / 1073741824.0, persist_bytes / 1073741824.0, temp_bytes / 1073741824.0);
        if (tape_units_possible > 0)
            fprintf(stderr, "[tne-plan] tape retention: %d of this rank's %d (segment, slice) units keep their forward tape (%.3f GiB), the others are recomputed\n",
                    tape_units, tape_units_possible, tape_bytes / 1073741824.0);
        for (size_t pi = 0; pi < phases.size(); ++pi) {
            const Phase& ph = phases[pi];
            if (ph.nsl <= 1 || ph.zero_at_start.empty()) continue;
            double by[3] = {0, 0, 0};
            int n[3] = {0, 0, 0};
            for (int v : ph.zero_at_start) {
                auto ci = ph.comm.find(v);
                const int k = ci == ph.comm.end() ? 0 : ci->second.kind;
                by[k] += 16.0 * (double)extent_of(P.vals[v]); n[k]++;
            }
            fprintf(stderr, "[tne-plan] phase %zu seg %d %s: checkpoints combined by all-reduce %d (%.2f GiB), reduce-scatter %d (%.2f GiB), all-gather %d (%.2f GiB)\n",
                    pi, ph.seg, ph.backward ? "bwd" : "fwd", n[0], by[0] / 1073741824.0, n[1], by[1] / 1073741824.0, n[2], by[2] / 1073741824.0);
        }
        if (ctx->opt_debug == 2) return;   // summary only
        for (size_t pi = 0; pi < phases.size(); ++pi) {
            const Phase& ph = phases[pi];
            for (auto& po : ph.ops) {
                const GettDesc& d = po.plan.d;
                double fl = po.plan.flops, by = po.plan.bytes;
                const char* kind = po.plan.kernel == 0 ? "generic" : "dmma";
                if (po.s2) { kind = "site2"; fl = po.s2_flops; by = po.s2_bytes; }
                else if (po.s2b) { kind = "site2b"; fl = po.s2_flops; by = po.s2_bytes; }
                else if (!po.extra.empty()) {
                    kind = "multi";
                    for (auto& ex : po.extra) { fl += ex.plan.flops; by += ex.plan.bytes - 16.0 * (double)ex.plan.c_nelem; }
                }
                fprintf(stderr, "[tne-final] phase %zu seg %d %s nsl %lld first_only %d kind %s kernel %d re %d bwdop %d acc %d M %lld N %lld K %lld B %lld flops %.6e bytes %.6e out %d",
                        pi, ph.seg, ph.backward ? "bwd" : "fwd", (long long)ph.nsl, (int)po.first_slice_only, kind, po.plan.kernel,
                        (int)po.op.recompute, (int)po.op.backward, (int)po.op.acc, d.Msz, d.Nsz, d.Ksz, d.Bsz, fl, by, po.op.out);
                auto pv = [&](const char* nm, int v) {
                    if (v < 0) return;
                    const PView& w = ph.views.at(v);
                    fprintf(stderr, " %s %d:%lld:%s", nm, v, (long long)w.nelem, w.temp ? "t" : (P.vals[v].ext ? "x" : "c"));
                };
                if (po.s2) { pv("A", po.s2_a); pv("B1", po.s2_b1); pv("B2", po.s2_b2); }
                else if (po.s2b) { pv("dC", po.s2b_dC); pv("B2", po.s2b_B2); pv("B1", po.s2b_B1); pv("A", po.s2b_A); pv("dB1", po.s2b_dB1); }
                else { pv("a", po.op.a); pv("b", po.op.b); for (auto& ex : po.extra) { pv("a", ex.a); pv("b", ex.b); } }
                pv("o", po.op.out);
                fprintf(stderr, "\n");
            }
        }
    }

    bool seg_shard = false;   // deal segment-local slices to the ranks; all-reduce the checkpoints of sliced segments

    std::set<int> pending;    // checkpoints whose all-gather is in flight on the communication stream

    int run() {
        const int world = seg_shard ? std::max(1, ctx->world) : 1, rank = seg_shard ? ctx->rank : 0;
        pending.clear();
        for (auto& ph : phases) {
            const bool sharded = world > 1 && ph.nsl > 1;
            for (int v : ph.zero_at_start) {
                if (pending.count(v)) { TNE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm_done, 0)); pending.clear(); }
                TNE_TRY(launch_fill_zero(ctx, base_ptr(ph, v), extent_of(P.vals[v])));
            }
            std::vector<int64_t> idx(ph.sliced.size(), 0);
            for (int64_t it = 0; it < ph.nsl; ++it) {
                if (sharded && it % world != rank) continue;
                int64_t rem = it;
                for (size_t s = 0; s < idx.size(); ++s) { idx[s] = rem % ph.ssz[s]; rem /= ph.ssz[s]; }
                const int64_t tbase = ph.unit_tape_base.empty() ? -1 : ph.unit_tape_base[it];
                auto ptr_of = [&](int v) -> cplx* {
                    if (v < 0) return ctx->one_dev;
                    cplx* p = base_ptr(ph, v, tbase);
                    const PView& pv = ph.views.at(v);
                    for (size_t s = 0; s < idx.size(); ++s) p += idx[s] * pv.sl_stride[s];
                    return p;
                };
                for (auto& po : ph.ops) {
                    if (po.first_slice_only && it > 0) continue;
                    if (po.skip_if_retained && tbase >= 0) continue;      // the unit kept this value from its forward phase
                    if (!pending.empty()) {
                        bool hit = pending.count(po.op.out) > 0 || (po.s2b && pending.count(po.s2b_dB1) > 0);
                        if (!hit) for (int v : po.inputs()) if (v >= 0 && pending.count(v)) { hit = true; break; }
                        if (hit) { TNE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm_done, 0)); pending.clear(); }
                    }
                    // replicated segment: every rank computes the same thing; only rank 0 may add to the outputs
                    if (world > 1 && !sharded && rank != 0 && P.vals[po.op.out].ext) continue;
                    if (po.s2) {
                        TNE_TRY(site2_launch(po, ptr_of(po.s2_a), ptr_of(po.s2_b1), ptr_of(po.s2_b2), ptr_of(po.op.out)));
                        if (po.op.recompute) P.re_flops += po.s2_flops; else { P.flops += po.s2_flops; P.bytes += po.s2_bytes; }
                        P.n_contract++;
                        continue;
                    }
                    if (po.s2b) {
                        // replicated segment on several ranks: only rank 0 may add to the (external) leaf gradient; the
                        // others still need dA and simply never reduce their partials
                        const bool skip_leaf = world > 1 && !sharded && rank != 0 && P.vals[po.s2b_dB1].ext;
                        TNE_TRY(site2b_launch(po, ptr_of(po.s2b_dC), ptr_of(po.s2b_B2), ptr_of(po.s2b_B1), ptr_of(po.s2b_A), ptr_of(po.op.out),
                                              ptr_of(po.s2b_dB1), skip_leaf));
                        P.flops += po.s2_flops; P.bytes += po.s2_bytes;
                        P.n_contract++;
                        continue;
                    }
                    if (po.extra.empty()) {
                        TNE_TRY(contract_run(ctx, po.plan, ptr_of(po.op.a), ptr_of(po.op.b), ptr_of(po.op.out)));
                    } else {
                        const ContractPlan* plans[TNE_MULTI_MAX];
                        const cplx *As[TNE_MULTI_MAX], *Bs[TNE_MULTI_MAX];
                        int T = 0;
                        plans[T] = &po.plan; As[T] = ptr_of(po.op.a); Bs[T] = ptr_of(po.op.b); ++T;
                        for (auto& ex : po.extra) { plans[T] = &ex.plan; As[T] = ptr_of(ex.a); Bs[T] = ptr_of(ex.b); ++T; }
                        TNE_TRY(contract_run_multi(ctx, T, plans, As, Bs, ptr_of(po.op.out)));
                        for (auto& ex : po.extra) {
                            if (po.op.recompute) P.re_flops += ex.plan.flops;
                            else { P.flops += ex.plan.flops; P.bytes += ex.plan.bytes - 16.0 * (double)ex.plan.c_nelem; }
                        }
                    }
                    if (po.op.recompute) P.re_flops += po.plan.flops; else { P.flops += po.plan.flops; P.bytes += po.plan.bytes; }
                    P.n_contract++;
                }
            }
            if (sharded && !ph.zero_at_start.empty()) {
                // All checkpoints of the column in ONE NCCL group (a single fused launch instead of one or two per tensor).
                // The all-gathers (cotangent checkpoints of the reverse pass: every rank wrote the slices it owns) are not
                // needed before the next column's reverse ops, which come after that column's recompute: they run on a second
                // stream, the compute stream waits for them right before the first op that touches one of the tensors.
                const bool overlap = ctx->opt_comm_overlap && ctx->opt_rs_comm;
                for (int pass = 0; pass < 2; ++pass) {          // pass 0: compute stream (all-reduce, reduce-scatter); pass 1: all-gathers
                    std::vector<int> todo;
                    for (int v : ph.zero_at_start) {
                        auto ci = ph.comm.find(v);
                        const int kind = ci == ph.comm.end() ? 0 : ci->second.kind;
                        if ((kind == 2) == (pass == 1)) todo.push_back(v);
                    }
                    if (todo.empty()) continue;
                    const bool side = pass == 1 && overlap;
                    if (side) {
                        if (!pending.empty()) { TNE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm_done, 0)); pending.clear(); }
                        if (!ctx->comm_stream) {
                            TNE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
                            TNE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm_ready, cudaEventDisableTiming));
                            TNE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm_done, cudaEventDisableTiming));
                        }
                        TNE_CUDA(ctx, cudaEventRecord(ctx->ev_comm_ready, ctx->stream));
                        TNE_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comm_ready, 0));
                        ctx->comm_active = ctx->comm_stream;
                    }
                    int rc = comm_group_begin(ctx);
                    for (int v : todo) {
                        if (rc) break;
                        auto ci = ph.comm.find(v);
                        const int kind = ci == ph.comm.end() ? 0 : ci->second.kind;
                        if (kind == 1) rc = comm_reduce_scatter_blocks(ctx, base_ptr(ph, v), ci->second.block, ci->second.nblocks);
                        else if (kind == 2) rc = comm_all_gather_blocks(ctx, base_ptr(ph, v), ci->second.block, ci->second.nblocks);
                        else rc = comm_allreduce_ptr(ctx, base_ptr(ph, v), extent_of(P.vals[v]));
                    }
                    const int rc2 = comm_group_end(ctx);
                    ctx->comm_active = nullptr;
                    if (rc) return rc;
                    TNE_TRY(rc2);
                    if (side) {
                        TNE_CUDA(ctx, cudaEventRecord(ctx->ev_comm_done, ctx->comm_stream));
                        pending.insert(todo.begin(), todo.end());
                    }
                }
            }
        }
        if (!pending.empty()) { TNE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm_done, 0)); pending.clear(); }
        return TNE_OK;
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------
// compiled programs and their cache
// ---------------------------------------------------------------------------------------
namespace {

// Everything tne_expect_terms builds on the host for one (tree, leaf shapes, term pattern, gradient request, options):
// the op lists, the phases with their classified launch plans and the memory plan.  Device pointers enter only through
// Val::ext of the external values, which are re-bound per call -- a TDVP driver evaluates the same program every step.
struct Compiled {
    Program P;
    Exec E;
    std::vector<std::vector<int32_t>> leaf_labels;
    std::vector<int32_t> sliced;
    int root = -1, seed_val = -1;
    std::vector<int> grad_vals;
    int64_t peak = 0;
    explicit Compiled(tne_ctx* c) : P(), E((P.ctx = c, P)) {}
};

template <typename T>
void key_put(std::string& k, const T* p, size_t n) { k.append(reinterpret_cast<const char*>(p), n * sizeof(T)); }
template <typename T>
void key_put1(std::string& k, T v) { key_put(k, &v, 1); }

// byte string that determines the compiled program
int program_key(tne_ctx* ctx, const tne_tree* T, const tne_buf* leaves, const int32_t* leaf_conj, int n_terms, const tne_term* terms,
                const std::map<int64_t, int>& repl_id, int n_grad, const int32_t* grad_leaves, std::string& k) {
    k.clear();
    key_put1(k, T->n_leaves); key_put1(k, T->n_nodes); key_put1(k, T->n_sliced); key_put1(k, T->shard_rank); key_put1(k, T->shard_world);
    key_put1(k, T->seg_shard); key_put1(k, T->n_segments);
    if (T->n_leaves <= 0 || T->n_nodes < 0) return tne_fail(ctx, TNE_ERR_INVALID, "empty tree");
    int nlab = 0;
    for (int i = 0; i < T->n_leaves; ++i) nlab += T->leaf_rank[i];
    key_put(k, T->leaf_rank, T->n_leaves); key_put(k, T->leaf_labels, nlab);
    key_put(k, T->node_children, 2 * (size_t)T->n_nodes); key_put(k, T->node_out_rank, T->n_nodes);
    int nout = 0;
    for (int j = 0; j < T->n_nodes; ++j) nout += T->node_out_rank[j];
    key_put(k, T->node_out_labels, nout);
    if (T->n_sliced > 0) key_put(k, T->sliced_labels, T->n_sliced);
    if (T->n_segments > 1 && T->seg_first_node) {
        key_put(k, T->seg_first_node, T->n_segments);
        if (T->seg_sliced_off) {
            key_put(k, T->seg_sliced_off, T->n_segments + 1);
            key_put(k, T->seg_sliced_labels, T->seg_sliced_off[T->n_segments]);
        }
    }
    for (int i = 0; i < T->n_leaves; ++i) {
        Buf* b = buf_get(ctx, leaves[i]);
        if (!b) return TNE_ERR_INVALID;
        key_put1(k, (int32_t)b->dims.size());
        key_put(k, b->dims.data(), b->dims.size());
        key_put1(k, (int32_t)(leaf_conj ? leaf_conj[i] != 0 : 0));
    }
    key_put1(k, n_terms);
    for (int t = 0; t < n_terms; ++t) {
        key_put1(k, terms[t].n_repl);
        for (int r = 0; r < terms[t].n_repl; ++r) { key_put1(k, terms[t].leaf[r]); key_put1(k, (int32_t)repl_id.at(terms[t].repl[r])); }
        key_put1(k, terms[t].coef_re); key_put1(k, terms[t].coef_im);
    }
    key_put1(k, n_grad);
    if (n_grad > 0) key_put(k, grad_leaves, n_grad);
    const int64_t opts[] = {ctx->opt_force_generic, ctx->opt_debug, ctx->opt_pad, ctx->opt_no_rows, ctx->opt_no_pdl, ctx->opt_no_site2, ctx->opt_no_fuse,
                            ctx->opt_no_permute, ctx->opt_no_gemm, ctx->opt_no_site2b, ctx->opt_absorb_tma, ctx->opt_no_reduce_tma, ctx->opt_rs_comm, ctx->opt_tape_gib,
                            (int64_t)ctx->world, (int64_t)ctx->rank, ctx->mem_limit};
    key_put(k, opts, sizeof(opts) / sizeof(opts[0]));
    return TNE_OK;
}

constexpr size_t PLAN_CACHE_MAX = 6;

}  // namespace

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" int tne_expect_terms(tne_ctx* ctx, const tne_tree* tree, const tne_buf* leaves, const int32_t* leaf_conj,
                                int n_terms, const tne_term* terms, int n_grad, const int32_t* grad_leaves,
                                double seed_re, double seed_im, tne_buf* value_out, tne_buf* grads_out) {
    TNE_ENTER(ctx);
    if (!tree || !leaves || !value_out) return tne_fail(ctx, TNE_ERR_INVALID, "null argument");
    if (n_grad > 0 && !grads_out) return tne_fail(ctx, TNE_ERR_INVALID, "grads_out is null");
    if (n_terms > 0 && !terms) return tne_fail(ctx, TNE_ERR_INVALID, "terms is null");
    // replacement tensors by order of first appearance: terms that share a handle share a pattern of the dynamic programme
    std::map<int64_t, int> repl_id;
    std::vector<tne_buf> repl_handles;
    for (int k = 0; k < n_terms; ++k) {
        if (terms[k].n_repl < 0 || terms[k].n_repl > 4) return tne_fail(ctx, TNE_ERR_INVALID, "term %d: n_repl must be 0..4", k);
        for (int r = 0; r < terms[k].n_repl; ++r)
            if (!repl_id.count(terms[k].repl[r])) { repl_id[terms[k].repl[r]] = (int)repl_handles.size(); repl_handles.push_back(terms[k].repl[r]); }
    }
    std::string key;
    TNE_TRY(program_key(ctx, tree, leaves, leaf_conj, n_terms, terms, repl_id, n_grad, grad_leaves, key));
    std::shared_ptr<Compiled> cp;
    const bool use_cache = !ctx->opt_no_plan_cache && !ctx->plan_only && ctx->opt_debug <= 2;
    if (use_cache)
        for (auto it = ctx->plan_cache.begin(); it != ctx->plan_cache.end(); ++it)
            if (it->first == key) {
                cp = std::static_pointer_cast<Compiled>(it->second);
                ctx->plan_cache.splice(ctx->plan_cache.begin(), ctx->plan_cache, it);   // most recently used first
                break;
            }
    // a cached programme planned its tape blocks against the memory that was free THEN: if that no longer fits, plan again
    if (cp && cp->E.tape_bytes > 0 && cp->peak > ctx->mem_limit - ctx->in_use) {
        for (auto it = ctx->plan_cache.begin(); it != ctx->plan_cache.end(); ++it)
            if (it->first == key) { ctx->plan_cache.erase(it); break; }
        cp.reset();
    }
    const bool fresh = !cp;
    if (fresh) cp = std::make_shared<Compiled>(ctx);
    Program& P = cp->P;
    Exec& E = cp->E;

    // outputs live in user-visible buffers, cleared once; every write into them accumulates
    // (over global slices, segment slices and repeated contributions alike)
    std::vector<tne_buf> owned;
    auto fail_cleanup = [&](int rc) { for (auto h : owned) tne_free(ctx, h); return rc; };
    std::vector<tne_buf> hgrads;
    tne_buf hval = 0, hseed = 0;

    if (fresh) {
        Builder B(ctx, P, tree);
        B.n_terms = n_terms;
        B.terms = terms;
        B.repl_id = repl_id;
        int root = -1;
        TNE_TRY(build_forward(B, leaves, leaf_conj, true, &root));
        P.root_val = root;
        cp->root = root;
        cp->leaf_labels = B.leaf_labels;
        cp->sliced = B.sliced;
        E.seg_shard = tree->seg_shard != 0 && ctx->world > 1;
        if (E.seg_shard && tree->n_sliced > 0)
            return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "seg_shard cannot be combined with globally sliced labels");
        TNE_TRY(E.assign_segments(tree));
        TNE_TRY(buf_new(ctx, P.vals[root].dims, &hval));
        owned.push_back(hval);
        P.vals[root].ext = buf_get(ctx, hval)->ptr();
        P.vals[root].bind_kind = 3;
        if (n_grad > 0) {
            if (P.vals[root].nelem != 1) return fail_cleanup(tne_fail(ctx, TNE_ERR_UNSUPPORTED, "gradients need a scalar output"));
            int rc = buf_new(ctx, {}, &hseed);
            if (rc) return fail_cleanup(rc);
            owned.push_back(hseed);
            int seed_val = P.new_val({}, {});
            P.vals[seed_val].ext = buf_get(ctx, hseed)->ptr();
            P.vals[seed_val].bind_kind = 4;
            P.written[seed_val] = true;
            cp->seed_val = seed_val;
            for (int g = 0; g < n_grad; ++g) {
                int lf = grad_leaves[g];
                if (lf < 0 || lf >= tree->n_leaves) return fail_cleanup(tne_fail(ctx, TNE_ERR_INVALID, "gradient requested for unknown leaf %d", lf));
                Buf* lb = buf_get(ctx, leaves[lf]);
                tne_buf hg;
                rc = buf_new(ctx, lb->dims, &hg);
                if (rc) return fail_cleanup(rc);
                owned.push_back(hg);
                hgrads.push_back(hg);
                int lv = B.nodes[lf].var[Key()];
                P.vals[lv].needs_grad = true;
                Val gv;
                gv.labels = P.vals[lv].labels;
                gv.dims = P.vals[lv].dims;
                gv.strides = P.vals[lv].strides;
                gv.nelem = P.vals[lv].nelem;
                gv.ext = buf_get(ctx, hg)->ptr();
                gv.bind_kind = 5; gv.bind_index = g;
                gv.leaf = lf;
                P.vals.push_back(gv);
                P.written.push_back(true);
                P.vals[lv].grad = (int)P.vals.size() - 1;
                cp->grad_vals.push_back(P.vals[lv].grad);
            }
            build_backward(P, root, seed_val);
        }
        {
            int rc = E.build(tree, n_grad > 0);
            if (rc) return fail_cleanup(rc);
        }
        {
            int w = E.seg_shard ? ctx->world : 1;
            if (ctx->plan_only && ctx->opt_rs_comm > 1) w = (int)ctx->opt_rs_comm;   // planner: rs_comm = the world size to plan for
            E.plan_comm(w);
        }
        {
            // memory left after the checkpoints and the per-segment arena keeps forward tapes (reverse mode only)
            int64_t tb = 0;
            if (n_grad > 0 && ctx->opt_tape_gib != 0) {
                const int64_t margin = (int64_t)6 << 30;
                tb = ctx->opt_tape_gib > 0 ? (ctx->opt_tape_gib << 30) : ctx->mem_limit - ctx->in_use - E.persist_bytes - E.temp_bytes - margin;
                if (ctx->plan_only && ctx->opt_tape_gib < 0) tb = ((int64_t)151 << 30) - E.persist_bytes - E.temp_bytes - margin;   // planner: a B200
            }
            const int w = E.seg_shard ? ctx->world : 1;
            E.plan_tape(tb, (ctx->plan_only && ctx->opt_rs_comm > 1) ? (int)ctx->opt_rs_comm : w, E.seg_shard ? ctx->rank : 0);
        }
        if (ctx->opt_debug > 2 || ctx->plan_only) E.dump_final();
        if (ctx->plan_only)
            return fail_cleanup(tne_fail(ctx, TNE_ERR_UNSUPPORTED, "planner context (TNE_PLAN_ONLY): program built and dumped, nothing computed"));
        cp->peak = E.persist_bytes + E.temp_bytes + E.tape_bytes;
        if (use_cache) {
            ctx->plan_cache.emplace_front(key, std::static_pointer_cast<void>(cp));
            while (ctx->plan_cache.size() > PLAN_CACHE_MAX) ctx->plan_cache.pop_back();
        }
    } else {
        // cached program: new output buffers, then every external value is re-bound to this call's buffers
        int rc = buf_new(ctx, P.vals[cp->root].dims, &hval);
        if (rc) return rc;
        owned.push_back(hval);
        if (n_grad > 0) {
            rc = buf_new(ctx, {}, &hseed);
            if (rc) return fail_cleanup(rc);
            owned.push_back(hseed);
            for (int g = 0; g < n_grad; ++g) {
                tne_buf hg;
                rc = buf_new(ctx, buf_get(ctx, leaves[grad_leaves[g]])->dims, &hg);
                if (rc) return fail_cleanup(rc);
                owned.push_back(hg);
                hgrads.push_back(hg);
            }
        }
        for (auto& v : P.vals) {
            switch (v.bind_kind) {
                case 1: v.ext = buf_get(ctx, leaves[v.bind_index])->ptr(); break;
                case 2: {
                    Buf* b = buf_get(ctx, repl_handles[v.bind_index]);
                    if (!b) return fail_cleanup(TNE_ERR_INVALID);
                    if (b->dims != buf_get(ctx, leaves[v.leaf])->dims)
                        return fail_cleanup(tne_fail(ctx, TNE_ERR_INVALID, "replacement for leaf %d has a different shape", v.leaf));
                    v.ext = b->ptr();
                    break;
                }
                case 3: v.ext = buf_get(ctx, hval)->ptr(); break;
                case 4: v.ext = buf_get(ctx, hseed)->ptr(); break;
                case 5: v.ext = buf_get(ctx, hgrads[v.bind_index])->ptr(); break;
                default: break;
            }
            v.slice_base = 0;
        }
        P.flops = P.bytes = P.re_flops = 0;
        P.n_contract = 0;
    }
    ctx->plan_cache_hits += fresh ? 0 : 1;
    const int root = cp->root;
    {
        Buf* vb = buf_get(ctx, hval);
        int rc = launch_fill_zero(ctx, vb->ptr(), vb->nelem);
        if (rc) return fail_cleanup(rc);
    }
    if (n_grad > 0) {
        // the cotangent seed is written by a kernel (no host synchronisation: planning of the next call overlaps this one)
        int rc = launch_fill_value(ctx, buf_get(ctx, hseed)->ptr(), 1, seed_re, seed_im);
        for (size_t g = 0; g < hgrads.size() && !rc; ++g) rc = launch_fill_zero(ctx, buf_get(ctx, hgrads[g])->ptr(), buf_get(ctx, hgrads[g])->nelem);
        if (rc) return fail_cleanup(rc);
    }
    const int64_t budget = ctx->mem_limit - ctx->in_use;
    const int64_t peak = cp->peak;
    if (peak > budget)
        return fail_cleanup(tne_fail(ctx, TNE_ERR_OOM,
                                     "tree program needs %lld bytes of workspace (%lld checkpoints + %lld per-segment) but the budget is %lld; "
                                     "use more segments / slice more labels",
                                     (long long)peak, (long long)E.persist_bytes, (long long)E.temp_bytes, (long long)budget));
    P.ws_bytes = peak;
    {
        int rc = ws_reserve(ctx, peak);
        if (rc) return fail_cleanup(rc);
    }

    // global slice loop (SlicedEinsum), sharded by rank
    std::vector<int64_t> ssz(tree->n_sliced, 1);
    for (int s = 0; s < tree->n_sliced; ++s) {
        bool found = false;
        for (int i = 0; i < tree->n_leaves && !found; ++i) {
            int q = find_in(cp->leaf_labels[i], tree->sliced_labels[s]);
            if (q >= 0) { ssz[s] = buf_get(ctx, leaves[i])->dims[q]; found = true; }
        }
        if (!found) return fail_cleanup(tne_fail(ctx, TNE_ERR_INVALID, "sliced label %d is not on any leaf", (int)tree->sliced_labels[s]));
    }
    int64_t nslices = 1;
    for (auto s : ssz) nslices *= s;
    const int world = std::max(1, (int)tree->shard_world), rank = tree->shard_rank;
    for (int64_t it = 0; it < nslices; ++it) {
        if (it % world != rank) continue;
        if (tree->n_sliced > 0) {
            std::vector<int64_t> idx(tree->n_sliced);
            int64_t rem = it;
            for (int s = 0; s < tree->n_sliced; ++s) { idx[s] = rem % ssz[s]; rem /= ssz[s]; }
            for (auto& v : P.vals) {
                if (v.leaf < 0) continue;
                const auto& full = cp->leaf_labels[v.leaf];
                Buf* lb = buf_get(ctx, leaves[v.leaf]);
                int64_t st = 1, base = 0;
                for (size_t q = 0; q < full.size(); ++q) {
                    int s = find_in(cp->sliced, full[q]);
                    if (s >= 0) base += idx[s] * st;
                    st *= lb->dims[q];
                }
                v.slice_base = base;
            }
        }
        int rc = E.run();
        if (rc) return fail_cleanup(rc);
    }
    if (E.seg_shard) {   // outputs are complete on every rank on return: one grouped NCCL call
        std::vector<cplx*> ptrs = {buf_get(ctx, hval)->ptr()};
        std::vector<int64_t> ns = {buf_get(ctx, hval)->nelem};
        for (auto h : hgrads) { ptrs.push_back(buf_get(ctx, h)->ptr()); ns.push_back(buf_get(ctx, h)->nelem); }
        int rc = comm_allreduce_many(ctx, (int)ptrs.size(), ptrs.data(), ns.data());
        if (rc) return fail_cleanup(rc);
    }
    ctx->stats[0] = P.flops; ctx->stats[1] = P.bytes; ctx->stats[2] = (double)P.n_contract;
    ctx->stats[3] = (double)peak; ctx->stats[4] = P.re_flops; ctx->stats[5] = (double)P.vals.size();
    if (ctx->opt_debug) {
        fprintf(stderr, "[tne] program%s: %zu fwd ops, %zu bwd ops, %zu phases, workspace %.3f GiB (checkpoints %.3f + segment %.3f), flops %.3e (+%.3e recomputed), %lld launches\n",
                fresh ? "" : " (cached plan)", P.fwd.size(), P.bwd.size(), E.phases.size(), peak / 1073741824.0, E.persist_bytes / 1073741824.0,
                E.temp_bytes / 1073741824.0, P.flops, P.re_flops, (long long)P.n_contract);
        if (ctx->opt_debug > 1)
            for (auto& ph : E.phases)
                fprintf(stderr, "[tne]   phase seg %d %s: %zu ops x %lld slices, temp %.3f GiB\n", ph.seg, ph.backward ? "bwd" : "fwd",
                        ph.ops.size(), (long long)ph.nsl, ph.temp_bytes / 1073741824.0);
    }
    if (hseed) tne_free(ctx, hseed);
    *value_out = hval;
    for (int g = 0; g < n_grad; ++g) grads_out[g] = hgrads[g];
    (void)root;
    return TNE_OK;
}

extern "C" int tne_contract_tree(tne_ctx* ctx, const tne_tree* tree, const tne_buf* leaves, const int32_t* leaf_conj, tne_buf* out) {
    return tne_expect_terms(ctx, tree, leaves, leaf_conj, 0, nullptr, 0, nullptr, 1.0, 0.0, out, nullptr);
}
