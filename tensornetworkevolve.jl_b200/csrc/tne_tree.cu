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
//   * memory: intermediates get static offsets in one workspace arena from exact lifetimes;
//     if the plan exceeds the budget, the largest forward values are dropped after their last
//     forward use and recomputed segment-wise when the backward pass needs them;
//   * slicing: sliced labels are looped over (sharded by rank), leaves become strided views.
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

TensorRef ref_of(const Program& P, int id, bool conj) {
    const Val& v = P.vals[id];
    TensorRef t;
    t.labels = v.labels;
    t.dims = v.dims;
    t.strides = v.strides;
    t.conj = conj;
    t.ptr = v.ext ? v.ext + v.slice_base : reinterpret_cast<cplx*>(reinterpret_cast<char*>(P.ctx->ws) + v.offset);
    return t;
}

// ---------------------------------------------------------------------------------------
// static memory plan: greedy-by-size offset assignment over exact lifetimes
// ---------------------------------------------------------------------------------------
int64_t plan_memory(Program& P) {
    const int nv = (int)P.vals.size();
    std::vector<int> first(nv, -1), last(nv, -1);
    for (int i = 0; i < (int)P.all.size(); ++i) {
        const Op& op = P.all[i];
        int ids[3] = {op.out, op.a, op.b};
        for (int id : ids) {
            if (id < 0 || P.vals[id].ext) continue;
            if (first[id] < 0) first[id] = i;
            last[id] = i;
        }
    }
    std::vector<int> order;
    for (int v = 0; v < nv; ++v) if (first[v] >= 0) order.push_back(v);
    auto bytes_of = [&](int v) { int64_t b = P.vals[v].nelem * (int64_t)sizeof(cplx); return (b + 511) / 512 * 512; };
    std::sort(order.begin(), order.end(), [&](int x, int y) {
        int64_t bx = bytes_of(x), by = bytes_of(y);
        if (bx != by) return bx > by;
        return first[x] < first[y];
    });
    std::vector<int> placed;
    int64_t peak = 0;
    for (int v : order) {
        std::vector<std::pair<int64_t, int64_t>> busy;
        for (int u : placed)
            if (!(last[u] < first[v] || last[v] < first[u])) busy.emplace_back(P.vals[u].offset, P.vals[u].offset + bytes_of(u));
        std::sort(busy.begin(), busy.end());
        int64_t off = 0, need = bytes_of(v);
        for (auto& iv : busy) {
            if (iv.first - off >= need) break;
            off = std::max(off, iv.second);
        }
        P.vals[v].offset = off;
        placed.push_back(v);
        peak = std::max(peak, off + need);
    }
    return peak;
}

// ---------------------------------------------------------------------------------------
// final op sequence with drop/recompute of forward values larger than `threshold` bytes
// ---------------------------------------------------------------------------------------
void schedule(Program& P, int64_t threshold) {
    P.all.clear();
    const int nv0 = (int)P.vals.size();
    std::vector<bool> used_bwd(nv0, false);
    for (auto& op : P.bwd) { if (op.a >= 0 && op.a < nv0) used_bwd[op.a] = true; if (op.b >= 0 && op.b < nv0) used_bwd[op.b] = true; }
    std::vector<std::vector<int>> producers(nv0);
    for (int i = 0; i < (int)P.fwd.size(); ++i) producers[P.fwd[i].out].push_back(i);
    std::vector<bool> dropped(nv0, false);
    for (int v = 0; v < nv0; ++v)
        if (!P.vals[v].ext && !producers[v].empty() && used_bwd[v] && threshold >= 0 &&
            P.vals[v].nelem * (int64_t)sizeof(cplx) > threshold && v != P.root_val)
            dropped[v] = true;
    std::vector<int> cur(nv0);
    for (int v = 0; v < nv0; ++v) cur[v] = v;
    std::vector<bool> alive(nv0, false);
    for (int v = 0; v < nv0; ++v) if (P.vals[v].ext) alive[v] = true;
    std::vector<int> last_fwd(nv0, -1);
    for (int i = 0; i < (int)P.fwd.size(); ++i) { if (P.fwd[i].a >= 0) last_fwd[P.fwd[i].a] = i; if (P.fwd[i].b >= 0) last_fwd[P.fwd[i].b] = i; }
    auto remap = [&](Op op) {
        if (op.out < nv0) op.out = cur[op.out];
        if (op.a >= 0 && op.a < nv0) op.a = cur[op.a];
        if (op.b >= 0 && op.b < nv0) op.b = cur[op.b];
        return op;
    };
    for (int i = 0; i < (int)P.fwd.size(); ++i) {
        const Op& op = P.fwd[i];
        P.all.push_back(remap(op));
        alive[op.out] = true;
        int ins[2] = {op.a, op.b};
        for (int v : ins) if (v >= 0 && dropped[v] && last_fwd[v] == i) alive[v] = false;
    }
    for (int v = 0; v < nv0; ++v) if (dropped[v] && last_fwd[v] < 0) alive[v] = false;  // never read forward
    std::function<void(int)> ensure = [&](int v) {
        if (v < 0 || v >= nv0 || alive[v] || producers[v].empty()) return;  // only forward values are recomputed
        for (int pi : producers[v]) { ensure(P.fwd[pi].a); ensure(P.fwd[pi].b); }
        // fresh instance so the two lifetimes get separate workspace intervals
        Val clone = P.vals[v];
        clone.offset = -1;
        P.vals.push_back(clone);
        P.written.push_back(true);
        cur[v] = (int)P.vals.size() - 1;
        for (int pi : producers[v]) {
            Op op = remap(P.fwd[pi]);
            op.recompute = true;
            P.all.push_back(op);
        }
        alive[v] = true;
    };
    for (auto& op0 : P.bwd) {
        ensure(op0.a);
        ensure(op0.b);
        P.all.push_back(remap(op0));
    }
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
    std::map<int64_t, int> repl_val;       // (leaf, handle) -> value id, keyed by handle*4096+leaf

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
            P.vals[v].conj = base.conj;
            P.vals[v].leaf = lf;
            P.written[v] = true;
            B.nodes[lf].var[key] = v;
        }
    }
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
        P.emit(P.fwd, op);
    }
    // internal nodes
    int lpos = 0;
    for (int j = 0; j < nn; ++j) {
        int ca = T->node_children[2 * j], cb = T->node_children[2 * j + 1];
        int self = nl + j;
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
            g.kind = OP_AXPY; g.backward = true;
            g.out = grad_of(op.a); g.a = dout;
            if (op.conj_a) { g.conj_a = true; g.are = op.are; g.aim = op.aim; }       // a' = conj(a): da = conj(conj(alpha) dout)
            else { g.conj_a = false; g.are = op.are; g.aim = -op.aim; }
            P.emit(P.bwd, g);
            continue;
        }
        // c (+)= alpha * fa(a) * fb(b)
        if (P.vals[op.a].needs_grad) {
            Op g;
            g.kind = OP_CONTRACT; g.backward = true;
            g.out = grad_of(op.a); g.a = dout; g.b = op.b;
            if (op.conj_a) { g.conj_a = true; g.conj_b = op.conj_b; g.are = op.are; g.aim = op.aim; }
            else { g.conj_a = false; g.conj_b = !op.conj_b; g.are = op.are; g.aim = -op.aim; }
            P.emit(P.bwd, g);
        }
        if (P.vals[op.b].needs_grad) {
            Op g;
            g.kind = OP_CONTRACT; g.backward = true;
            g.out = grad_of(op.b); g.a = op.a; g.b = dout;
            if (op.conj_b) { g.conj_b = true; g.conj_a = op.conj_a; g.are = op.are; g.aim = op.aim; }
            else { g.conj_b = false; g.conj_a = !op.conj_a; g.are = op.are; g.aim = -op.aim; }
            P.emit(P.bwd, g);
        }
    }
}

int run_ops(Program& P, bool count_stats) {
    tne_ctx* ctx = P.ctx;
    for (const Op& op : P.all) {
        TensorRef C = ref_of(P, op.out, false);
        if (op.kind == OP_AXPY) {
            TensorRef A = ref_of(P, op.a, op.conj_a);
            TensorRef one;
            one.ptr = ctx->one_dev;
            TNE_TRY(contract_launch(ctx, A, one, C, op.acc, op.are, op.aim));
            continue;
        }
        TensorRef A = ref_of(P, op.a, op.conj_a);
        TensorRef Bt = ref_of(P, op.b, op.conj_b);
        TNE_TRY(contract_launch(ctx, A, Bt, C, op.acc, op.are, op.aim));
        if (count_stats) {
            double f, b;
            contract_cost(A, Bt, C, &f, &b);
            if (op.recompute) P.re_flops += f; else { P.flops += f; P.bytes += b; }
            P.n_contract++;
        }
    }
    return TNE_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" int tne_expect_terms(tne_ctx* ctx, const tne_tree* tree, const tne_buf* leaves, const int32_t* leaf_conj,
                                int n_terms, const tne_term* terms, int n_grad, const int32_t* grad_leaves,
                                double seed_re, double seed_im, tne_buf* value_out, tne_buf* grads_out) {
    if (!tree || !leaves || !value_out) return tne_fail(ctx, TNE_ERR_INVALID, "null argument");
    if (n_grad > 0 && !grads_out) return tne_fail(ctx, TNE_ERR_INVALID, "grads_out is null");
    Program P;
    P.ctx = ctx;
    Builder B(ctx, P, tree);
    B.n_terms = n_terms;
    B.terms = terms;
    int root = -1;
    TNE_TRY(build_forward(B, leaves, leaf_conj, true, &root));
    P.root_val = root;

    // output value (accumulated over slices) lives in a user-visible buffer
    tne_buf hval;
    TNE_TRY(buf_new(ctx, P.vals[root].dims, &hval));
    const bool sliced = tree->n_sliced > 0;
    std::vector<tne_buf> hgrads;
    std::vector<int> grad_vals;
    int seed_val = -1;
    tne_buf hseed = 0;
    if (n_grad > 0) {
        if (P.vals[root].nelem != 1) { tne_free(ctx, hval); return tne_fail(ctx, TNE_ERR_UNSUPPORTED, "gradients need a scalar output"); }
        TNE_TRY(buf_new(ctx, {}, &hseed));
        cplx s = make_double2(seed_re, seed_im);
        TNE_CUDA(ctx, cudaMemcpyAsync(buf_get(ctx, hseed)->ptr(), &s, sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
        TNE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        seed_val = P.new_val({}, {});
        P.vals[seed_val].ext = buf_get(ctx, hseed)->ptr();
        P.written[seed_val] = true;
        for (int g = 0; g < n_grad; ++g) {
            int lf = grad_leaves[g];
            if (lf < 0 || lf >= tree->n_leaves) return tne_fail(ctx, TNE_ERR_INVALID, "gradient requested for unknown leaf %d", lf);
            Buf* lb = buf_get(ctx, leaves[lf]);
            tne_buf hg;
            TNE_TRY(buf_new(ctx, lb->dims, &hg));
            Buf* gb = buf_get(ctx, hg);
            TNE_TRY(launch_fill_zero(ctx, gb->ptr(), gb->nelem));
            hgrads.push_back(hg);
            int lv = B.nodes[lf].var[Key()];
            P.vals[lv].needs_grad = true;
            Val gv;
            gv.labels = P.vals[lv].labels;
            gv.dims = P.vals[lv].dims;
            gv.strides = P.vals[lv].strides;
            gv.nelem = P.vals[lv].nelem;
            gv.ext = gb->ptr();
            gv.leaf = lf;
            P.vals.push_back(gv);
            P.written.push_back(true);  // pre-zeroed: every write accumulates
            P.vals[lv].grad = (int)P.vals.size() - 1;
            grad_vals.push_back(P.vals[lv].grad);
        }
        build_backward(P, root, seed_val);
    }
    // root: write into the output buffer; with slices it accumulates across slices
    {
        Buf* vb = buf_get(ctx, hval);
        P.vals[root].ext = vb->ptr();
        if (sliced) {
            TNE_TRY(launch_fill_zero(ctx, vb->ptr(), vb->nelem));
            for (auto& op : P.fwd) if (op.out == root) op.acc = true;
        }
    }

    // memory plan with automatic recompute threshold
    int64_t budget = ctx->mem_limit - ctx->in_use;
    std::vector<int64_t> cands;
    for (auto& v : P.vals) if (!v.ext) cands.push_back(v.nelem * (int64_t)sizeof(cplx));
    std::sort(cands.begin(), cands.end(), std::greater<int64_t>());
    cands.erase(std::unique(cands.begin(), cands.end()), cands.end());
    const size_t nvals0 = P.vals.size();
    int64_t peak = 0;
    int64_t thr = ctx->opt_recompute_threshold;
    bool ok = false;
    if (thr >= 0 || n_grad == 0) {
        schedule(P, n_grad == 0 ? -1 : thr);
        peak = plan_memory(P);
        ok = peak <= budget;
    } else {
        schedule(P, -1);
        peak = plan_memory(P);
        ok = peak <= budget;
        for (size_t c = 0; !ok && c < cands.size(); ++c) {
            P.vals.resize(nvals0);
            P.written.resize(nvals0);
            schedule(P, cands[c] - 1);
            peak = plan_memory(P);
            ok = peak <= budget;
            thr = cands[c] - 1;
        }
    }
    if (!ok) {
        tne_free(ctx, hval);
        for (auto h : hgrads) tne_free(ctx, h);
        if (hseed) tne_free(ctx, hseed);
        return tne_fail(ctx, TNE_ERR_OOM, "tree program needs %lld bytes of workspace but the budget is %lld; slice more labels",
                        (long long)peak, (long long)budget);
    }
    P.ws_bytes = peak;
    TNE_TRY(ws_reserve(ctx, peak));

    // slice loop
    std::vector<int64_t> ssz(tree->n_sliced, 1);
    for (int s = 0; s < tree->n_sliced; ++s) {
        bool found = false;
        for (int i = 0; i < tree->n_leaves && !found; ++i) {
            int q = find_in(B.leaf_labels[i], tree->sliced_labels[s]);
            if (q >= 0) { ssz[s] = buf_get(ctx, leaves[i])->dims[q]; found = true; }
        }
        if (!found) return tne_fail(ctx, TNE_ERR_INVALID, "sliced label %d is not on any leaf", (int)tree->sliced_labels[s]);
    }
    int64_t nslices = 1;
    for (auto s : ssz) nslices *= s;
    const int world = std::max(1, (int)tree->shard_world), rank = tree->shard_rank;
    bool first = true;
    (void)first;
    for (int64_t it = 0; it < nslices; ++it) {
        if (it % world != rank) continue;
        if (tree->n_sliced > 0) {
            std::vector<int64_t> idx(tree->n_sliced);
            int64_t rem = it;
            for (int s = 0; s < tree->n_sliced; ++s) { idx[s] = rem % ssz[s]; rem /= ssz[s]; }
            for (auto& v : P.vals) {
                if (v.leaf < 0) continue;
                const auto& full = B.leaf_labels[v.leaf];
                Buf* lb = buf_get(ctx, leaves[v.leaf]);
                int64_t st = 1, base = 0;
                for (size_t q = 0; q < full.size(); ++q) {
                    int s = find_in(B.sliced, full[q]);
                    if (s >= 0) base += idx[s] * st;
                    st *= lb->dims[q];
                }
                v.slice_base = base;
            }
        }
        TNE_TRY(run_ops(P, true));
        first = false;
    }
    ctx->stats[0] = P.flops; ctx->stats[1] = P.bytes; ctx->stats[2] = (double)P.n_contract;
    ctx->stats[3] = (double)peak; ctx->stats[4] = P.re_flops; ctx->stats[5] = (double)P.vals.size();
    if (ctx->opt_debug)
        fprintf(stderr, "[tne] program: %zu fwd ops, %zu bwd ops, %zu scheduled, workspace %.3f GiB, recompute threshold %lld, flops %.3e (+%.3e recomputed)\n",
                P.fwd.size(), P.bwd.size(), P.all.size(), peak / 1073741824.0, (long long)thr, P.flops, P.re_flops);
    if (hseed) tne_free(ctx, hseed);
    *value_out = hval;
    for (int g = 0; g < n_grad; ++g) grads_out[g] = hgrads[g];
    return TNE_OK;
}

extern "C" int tne_contract_tree(tne_ctx* ctx, const tne_tree* tree, const tne_buf* leaves, const int32_t* leaf_conj, tne_buf* out) {
    return tne_expect_terms(ctx, tree, leaves, leaf_conj, 0, nullptr, 0, nullptr, 1.0, 0.0, out, nullptr);
}
