// voices.cpp — isomorphic-voice detection on a flat reference-style graph (see graph.hpp VoiceDetection; pure C++, no CUDA).
//
// Ports are ordered, so a voice has ONE canonical form: walk backwards from the channels it hands to the bus tree, channel by
// channel, and through every node's inputs in port order; number the nodes in the order they are first met. Two voices are
// isomorphic iff their canonical forms — node kind, port counts, the parameters that are not per-voice tables, and the canonical
// index / port (or graph_in channel, relative to the voice) behind every input — are equal. No search is involved.
#include <algorithm>
#include <map>

#include "graph.hpp"

namespace fw {
namespace {

struct Ep { Id node; uint32_t port = 0; bool ok = false; };  // what feeds an input port

struct Flat {
    Graph& g;
    std::unordered_map<uint64_t, std::vector<Ep>> in_of;                      // node -> producer per input port
    std::unordered_map<uint64_t, std::vector<std::vector<Id>>> cons_of;        // node -> consumers per output port
    explicit Flat(Graph& gr) : g(gr) {
        g.each_node([&](Id id, NodeRec& r) { in_of[id.pack()].assign(r.num_inputs, Ep{}); cons_of[id.pack()].assign(r.num_outputs, {}); });
        g.each_edge([&](Id, EdgeRec& e) {
            auto& ins = in_of[e.dst.pack()];
            if (e.dst_port < ins.size()) ins[e.dst_port] = Ep{e.src, e.src_port, true};
            auto& outs = cons_of[e.src.pack()];
            if (e.src_port < outs.size()) outs[e.src_port].push_back(e.dst);
        });
    }
    const NodeRec& rec(Id id) { return *g.node(id); }
    // a SumNode that can be a level of the bus tree over C channels: C outputs, `ports` groups of C inputs, every input connected
    bool tree_sum(Id id, uint32_t C, uint32_t* ports) {
        if (id == g.graph_in() || id == g.graph_out()) return false;
        const NodeRec& r = rec(id);
        if (r.params->kind != FW_NODE_SUM || r.num_outputs != C || (r.num_inputs != C && r.num_inputs != 2 * C)) return false;
        for (const Ep& e : in_of[id.pack()]) if (!e.ok) return false;
        *ports = r.num_inputs / C;
        return true;
    }
    // the C endpoints [first, first + C) of `dst` come from outputs 0..C-1 of ONE node, in order, and nothing else reads them
    bool whole_node_behind(Id dst, uint32_t first, uint32_t C, Id* src) {
        const auto& ins = in_of[dst.pack()];
        for (uint32_t c = 0; c < C; ++c) {
            const Ep& e = ins[first + c];
            if (!e.ok || e.port != c || (c > 0 && e.node != ins[first].node)) return false;
            if (cons_of[e.node.pack()][c].size() != 1) return false;
        }
        *src = ins[first].node;
        return true;
    }
};

// parameters that are part of a node's identity (everything that is not a per-voice table)
bool same_static(const NodeParams& a, const NodeParams& b) {
    if (a.kind != b.kind) return false;
    switch (a.kind) {
        case FW_NODE_HARD_CLIP: return a.threshold_gain == b.threshold_gain;
        case FW_NODE_BIQUAD: case FW_NODE_SVF: return a.num_stages == b.num_stages;
        case FW_NODE_DELAY: return a.delay == b.delay;
        case FW_NODE_CONV_REVERB: return a.ir_len == b.ir_len && a.ir_channels == b.ir_channels && a.ir == b.ir;
        case FW_NODE_RESAMPLER: return a.rs_phases == b.rs_phases && a.rs_taps == b.rs_taps && a.rs_table == b.rs_table;
        default: return true;
    }
}

struct Canon {
    std::vector<Id> order;                                        // canonical node order of one voice
    std::vector<std::vector<VoiceDetection::Src>> inputs;         // per node, per input port
    std::vector<VoiceDetection::Src> outputs;                     // per voice output channel
    std::vector<uint32_t> gin_channels;                           // absolute graph_in ports read, in first-met order
};

// canonical form of the voice that hands `leaf` (C endpoints) to the bus tree; false: it reaches back into `claimed` (another voice
// or the tree) or is malformed
bool canonical(Flat& f, const std::vector<Ep>& leaf, const std::unordered_map<uint64_t, uint32_t>& claimed, uint32_t voice, Canon* out, std::string* why) {
    std::unordered_map<uint64_t, int> index;
    std::vector<std::pair<Id, uint32_t>> stack;  // iterative DFS: (node, next input port)
    auto visit = [&](Id start) -> bool {
        if (index.count(start.pack())) return true;
        stack.push_back({start, 0});
        index[start.pack()] = (int)out->order.size(); out->order.push_back(start); out->inputs.emplace_back();
        while (!stack.empty()) {
            Id n = stack.back().first; uint32_t& port = stack.back().second;
            const auto& ins = f.in_of[n.pack()];
            if (port == 0) out->inputs[(size_t)index[n.pack()]].assign(ins.size(), VoiceDetection::Src{});
            if (port >= ins.size()) { stack.pop_back(); continue; }
            const Ep e = ins[port]; const uint32_t this_port = port++;
            VoiceDetection::Src& dst = out->inputs[(size_t)index[n.pack()]][this_port];
            if (!e.ok) { dst.node = -1; continue; }
            if (e.node == f.g.graph_in()) { dst.node = -2; dst.port = e.port; continue; }  // made voice-relative by the caller
            auto cl = claimed.find(e.node.pack());
            if (cl != claimed.end() && cl->second != voice) { *why = "voices share a node (voice " + std::to_string(voice) + " reaches into voice " + std::to_string(cl->second) + ")"; return false; }
            auto it = index.find(e.node.pack());
            if (it == index.end()) {
                const int idx = (int)out->order.size();
                index[e.node.pack()] = idx; out->order.push_back(e.node); out->inputs.emplace_back();
                dst.node = idx; dst.port = e.port;
                stack.push_back({e.node, 0});
            } else { dst.node = it->second; dst.port = e.port; }
        }
        return true;
    };
    for (const Ep& e : leaf) {
        VoiceDetection::Src s;
        if (!e.ok) s.node = -1;
        else if (e.node == f.g.graph_in()) { s.node = -2; s.port = e.port; }
        else {
            auto cl = claimed.find(e.node.pack());
            if (cl != claimed.end() && cl->second != voice) { *why = "voices share a node"; return false; }
            if (!visit(e.node)) return false;
            s.node = index[e.node.pack()]; s.port = e.port;
        }
        out->outputs.push_back(s);
    }
    return true;
}

}  // namespace

bool detect_voices(Graph& g, VoiceDetection* out, std::string* why) {
    *out = VoiceDetection{};
    Flat f(g);
    const NodeRec& gout = f.rec(g.graph_out());
    const uint32_t C = gout.num_inputs, n_gin = f.rec(g.graph_in()).num_outputs;
    if (C == 0) { *why = "graph_out has no inputs"; return false; }
    bool has_custom = false;
    g.each_node([&](Id, NodeRec& r) { if (r.params->kind == FW_NODE_CUSTOM) has_custom = true; });

    // ---- candidate bus trees, deepest first: levels[d] = the tree's nodes at depth d, left to right; the leaves are what feeds level D ----
    std::vector<std::vector<Id>> levels;  // levels[0] = {root}
    bool mixed_depth = false;
    {
        Id root; uint32_t ports = 0;
        if (f.whole_node_behind(g.graph_out(), 0, C, &root) && f.tree_sum(root, C, &ports)) {
            levels.push_back({root});
            for (;;) {  // one more level while EVERY child of the current level is a tree-eligible SumNode
                std::vector<Id> next; bool all = true;
                for (Id n : levels.back()) {
                    uint32_t np = 0; f.tree_sum(n, C, &np);
                    for (uint32_t q = 0; q < np && all; ++q) { Id ch; uint32_t cp = 0; if (f.whole_node_behind(n, q * C, C, &ch) && f.tree_sum(ch, C, &cp)) next.push_back(ch); else all = false; }
                    if (!all) break;
                }
                if (!all && !next.empty()) mixed_depth = true;  // some children of this level are tree SumNodes, some are not
                if (!all || next.empty()) break;
                levels.push_back(std::move(next));
            }
        }
    }
    // try depth D = levels.size() (leaves below the deepest level) down to 0 (no tree: one voice)
    std::string last_why; bool have_why = false;
    if (mixed_depth) { last_why = "the SumNode tree is not the balanced pairwise tree: its leaves are not all at the same depth (an unpaired last element must pass a 1-port SumNode)"; have_why = true; }
    for (size_t D = levels.size();; --D) {
        // leaves: the input groups of level D - 1, or graph_out's inputs for D == 0
        std::vector<std::vector<Ep>> leaves;
        if (D == 0) leaves.push_back(f.in_of[g.graph_out().pack()]);
        else for (Id n : levels[D - 1]) {
            uint32_t np = 0; f.tree_sum(n, C, &np);
            const auto& ins = f.in_of[n.pack()];
            for (uint32_t q = 0; q < np; ++q) leaves.emplace_back(ins.begin() + q * C, ins.begin() + (q + 1) * C);
        }
        bool ok = true; std::string w;
        // canonical shape: going up from the leaves, every node pairs two children except the last one of a level with an odd child count
        size_t n_child = leaves.size();
        for (size_t d = D; ok && d-- > 0;) {
            const auto& lv = levels[d];
            if (lv.size() != (n_child + 1) / 2) { ok = false; w = "the SumNode tree is not the balanced pairwise tree"; break; }
            for (size_t k = 0; k < lv.size(); ++k) {
                uint32_t np = 0; f.tree_sum(lv[k], C, &np);
                const uint32_t want = (k + 1 == lv.size() && (n_child & 1u)) ? 1u : 2u;
                if (np != want) { ok = false; w = "the SumNode tree is not the balanced pairwise tree (an unpaired last element must pass a 1-port SumNode)"; break; }
            }
            n_child = lv.size();
        }
        const uint32_t V = (uint32_t)leaves.size();
        if (ok && V > 1 && has_custom) { ok = false; w = "user nodes are one object per voice in a flat graph: batch them through fw_graph_add_custom_node on a batched context"; }
        if (ok && V > 1 && n_gin % V != 0) { ok = false; w = "graph_in's channels do not divide among the voices"; }
        VoiceDetection det; std::vector<Canon> cans(V);
        if (ok) {
            std::unordered_map<uint64_t, uint32_t> claimed;  // node -> voice; tree nodes -> UINT32_MAX
            for (size_t d = 0; d < D; ++d) for (Id n : levels[d]) { claimed[n.pack()] = UINT32_MAX; det.tree.push_back(n); }
            for (uint32_t v = 0; ok && v < V; ++v) {
                if (!canonical(f, leaves[v], claimed, v, &cans[v], &w)) { ok = false; break; }
                for (Id n : cans[v].order) claimed[n.pack()] = v;
            }
            // nothing of a voice may be read outside that voice (the tree reads exactly the leaf channels, checked through fan-out 1 below)
            for (uint32_t v = 0; ok && v < V; ++v) for (Id n : cans[v].order) for (const auto& port_cons : f.cons_of[n.pack()]) for (Id cns : port_cons) {
                auto it = claimed.find(cns.pack());
                const bool into_tree = it != claimed.end() && it->second == UINT32_MAX, gout_direct = cns == g.graph_out() && D == 0;
                if (!(into_tree || gout_direct || (it != claimed.end() && it->second == v))) { ok = false; w = "a voice feeds a node outside itself"; }
            }
            const uint32_t cin = V ? n_gin / V : 0;
            for (uint32_t v = 0; ok && v < V; ++v) {
                Canon& cn = cans[v];
                auto rel = [&](VoiceDetection::Src& s) { if (s.node != -2) return; if (s.port / (cin ? cin : 1u) != v || cin == 0) { ok = false; w = "voice " + std::to_string(v) + " reads graph_in channels outside its own range"; } else s.port -= v * cin; };
                for (auto& ports : cn.inputs) for (auto& s : ports) rel(s);
                for (auto& s : cn.outputs) rel(s);
                if (V > 1) for (size_t c = 0; ok && c < cn.outputs.size(); ++c) {  // the tree must be the only reader of a leaf channel
                    const Ep& e = leaves[v][c];
                    if (e.ok && e.node != g.graph_in() && f.cons_of[e.node.pack()][e.port].size() != 1) { ok = false; w = "a voice output is read by more than the bus tree"; }
                }
            }
            // isomorphism: every voice's canonical form equals voice 0's
            for (uint32_t v = 1; ok && v < V; ++v) {
                const Canon &a = cans[0], &b = cans[v];
                bool same = a.order.size() == b.order.size() && a.outputs.size() == b.outputs.size();
                for (size_t i = 0; same && i < a.order.size(); ++i) {
                    const NodeRec &ra = f.rec(a.order[i]), &rb = f.rec(b.order[i]);
                    same = ra.num_inputs == rb.num_inputs && ra.num_outputs == rb.num_outputs && same_static(*ra.params, *rb.params) && a.inputs[i].size() == b.inputs[i].size();
                    for (size_t p = 0; same && p < a.inputs[i].size(); ++p) same = a.inputs[i][p].node == b.inputs[i][p].node && a.inputs[i][p].port == b.inputs[i][p].port;
                }
                for (size_t c = 0; same && c < a.outputs.size(); ++c) same = a.outputs[c].node == b.outputs[c].node && a.outputs[c].port == b.outputs[c].port;
                if (!same) { ok = false; w = "voice " + std::to_string(v) + " is not isomorphic to voice 0"; }
            }
        }
        if (ok) {
            det.num_voices = V; det.voice_inputs = V ? n_gin / V : 0; det.voice_outputs = C;
            det.nodes.assign(cans[0].order.size(), std::vector<Id>(V));
            for (uint32_t v = 0; v < V; ++v) for (size_t i = 0; i < cans[v].order.size(); ++i) det.nodes[i][v] = cans[v].order[i];
            det.inputs = cans[0].inputs; det.outputs = cans[0].outputs;
            *out = std::move(det);
            if (D < levels.size()) *why = last_why;  // why the deeper reading of the tree was not taken (informational)
            return true;
        }
        if (!have_why) { last_why = w; have_why = true; }  // the deepest reading's reason is the informative one
        if (D == 0) break;
    }
    *why = last_why;
    return false;
}

}  // namespace fw
