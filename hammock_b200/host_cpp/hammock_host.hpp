// hammock_host.hpp -- C++ host side above the C ABI (include/hammock_b200.h).
//
// The reference is compiled code (Java) whose toolchain is absent from this image, so the host
// layer that a Hammock maintainer would keep in Java is mirrored here in C++ for the greedy path:
// same names, argument meaning and error behaviour as the reference classes (file:line cited,
// relative to /root/reference/src/cz/krejciadam/hammock/).  Parsing, ordering, automatic
// parameters and the result files live here; all scoring and clustering happens on the GPU
// behind hmk_greedy_cluster().  There is no CPU fallback.
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hammock_b200.h"

namespace hammock {

struct HammockException : std::runtime_error { using std::runtime_error::runtime_error; };      // HammockException.java
struct DataException : HammockException { using HammockException::HammockException; };          // DataException.java
struct FileFormatException : HammockException { using HammockException::HammockException; };    // FileFormatException.java
struct CLIException : HammockException { using HammockException::HammockException; };           // CLIException.java
struct NullPointerException : HammockException {   // LimitedGreedySequenceClusterer.java:104,108
    int step;
    explicit NullPointerException(int s)
        : HammockException("NullPointerException in greedy phase 1, step " + std::to_string(s)), step(s) {}
};
struct CudaException : HammockException { using HammockException::HammockException; };

static const char* const ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*";   // UniqueSequence.java:23-26

inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }   // Java int

// ---------------------------------------------------------------- Java collection semantics
inline int32_t java_string_hash(const std::string& s) {   // String.hashCode
    uint32_t h = 0;
    for (unsigned char c : s) h = 31u * h + c;
    return (int32_t)h;
}

// Iteration order of a java.util.HashMap<String, ?> (Java 8+: tail insertion, order-preserving
// resize) filled with `keys` in this order.  Bins are assumed not to treeify (< 8 collisions).
inline std::vector<std::string> java_hashmap_order(const std::vector<std::string>& keys) {
    size_t cap = 16;
    while (keys.size() > cap * 3 / 4) cap *= 2;
    std::vector<std::vector<std::string>> bins(cap);
    for (const auto& k : keys) {
        uint32_t h = (uint32_t)java_string_hash(k);
        h ^= h >> 16;
        bins[h & (cap - 1)].push_back(k);
    }
    std::vector<std::string> out;
    for (auto& b : bins) out.insert(out.end(), b.begin(), b.end());
    return out;
}

struct JavaRandom {   // java.util.Random
    uint64_t seed;
    explicit JavaRandom(int64_t s) : seed(((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1)) {}
    int32_t next(int bits) {
        seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)(int64_t)(seed >> (48 - bits));
    }
    int32_t nextInt(int32_t bound) {
        int32_t r = next(31), m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        for (int32_t u = r; (int32_t)((uint32_t)u - (uint32_t)(r = u % bound) + (uint32_t)m) < 0; u = next(31)) {}
        return r;
    }
};

// ---------------------------------------------------------------- data model
struct UniqueSequence {   // UniqueSequence.java:19-57, 81-109
    std::string sequence;                                  // upper-case letters (getSequenceString)
    std::vector<uint8_t> codes;                            // getSequence()
    std::vector<std::pair<std::string, int32_t>> labels;   // labelsMap (insertion order kept for determinism)

    UniqueSequence(const std::string& s, std::vector<std::pair<std::string, int32_t>> lab) : labels(std::move(lab)) {
        for (char c : s) {
            char u = (char)std::toupper((unsigned char)c);
            const char* p = u ? std::strchr(ALPHABET, u) : nullptr;
            if (!p) throw FileFormatException(std::string("Error, character ") + c +
                                              " is not a valid letter from the amino acid alphabet code.");   // :51-54
            codes.push_back((uint8_t)(p - ALPHABET));
            sequence.push_back(u);
        }
    }
    int32_t size() const {   // :81-88
        int32_t s = 0;
        for (auto& kv : labels) s = wrap_add(s, kv.second);
        return s;
    }
    int32_t count(const std::string& label) const {
        for (auto& kv : labels) if (kv.first == label) return kv.second;
        return 0;
    }
    bool has(const std::string& label) const {
        for (auto& kv : labels) if (kv.first == label) return true;
        return false;
    }
};

struct Cluster {   // Cluster.java:21-74, 113-123, 156-158
    int32_t id = 0;
    int32_t sizeSum = 0;
    std::vector<int> members;   // indices into the clustering-ordered sequence vector, insertion order
    int32_t getId() const { return id; }
    int32_t size() const { return sizeSum; }
    size_t getUniqueSize() const { return members.size(); }
};

// ---------------------------------------------------------------- loaders
inline bool java_ws(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == 0x0B || c == '\f' || c == '\r'; }

inline std::vector<std::string> read_lines(const std::string& path) {   // BufferedReader.readLine
    std::ifstream f(path, std::ios::binary);
    if (!f) throw HammockException("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string text = ss.str();
    std::vector<std::string> lines;
    size_t i = 0;
    while (i < text.size()) {
        size_t j = text.find('\n', i);
        if (j == std::string::npos) j = text.size();
        std::string l = text.substr(i, j - i);
        if (!l.empty() && l.back() == '\r') l.pop_back();
        lines.push_back(l);
        i = j + 1;
    }
    return lines;
}

inline int32_t parse_int(const std::string& tok) {   // Integer.parseInt
    size_t i = 0;
    bool neg = false;
    if (!tok.empty() && (tok[0] == '-' || tok[0] == '+')) { neg = tok[0] == '-'; i = 1; }
    if (i >= tok.size()) throw FileFormatException("NumberFormatException: For input string: \"" + tok + "\"");
    int64_t v = 0;
    for (; i < tok.size(); i++) {
        if (tok[i] < '0' || tok[i] > '9') throw FileFormatException("NumberFormatException: For input string: \"" + tok + "\"");
        v = v * 10 + (tok[i] - '0');
        if (v > (int64_t)INT32_MAX + 1) throw FileFormatException("NumberFormatException: For input string: \"" + tok + "\"");
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) throw FileFormatException("NumberFormatException: For input string: \"" + tok + "\"");
    return (int32_t)v;
}

inline int32_t decode_int(const std::string& tok) {   // Integer.decode
    std::string s = tok;
    bool neg = false;
    if (!s.empty() && s[0] == '-') { neg = true; s.erase(0, 1); } else if (!s.empty() && s[0] == '+') s.erase(0, 1);
    int radix = 10;
    if (s.size() >= 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) { radix = 16; s.erase(0, 2); }
    else if (!s.empty() && s[0] == '#') { radix = 16; s.erase(0, 1); }
    else if (s.size() > 1 && s[0] == '0') { radix = 8; s.erase(0, 1); }
    if (s.empty()) throw FileFormatException("NumberFormatException: For input string: \"" + tok + "\"");
    int64_t v = 0;
    for (char c : s) {
        int d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : 99;
        if (d >= radix) throw FileFormatException("NumberFormatException: For input string: \"" + tok + "\"");
        v = v * radix + d;
        if (v > (int64_t)INT32_MAX + 1) throw FileFormatException("NumberFormatException: For input string: \"" + tok + "\"");
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) throw FileFormatException("NumberFormatException: For input string: \"" + tok + "\"");
    return (int32_t)v;
}

inline std::string java_trim(const std::string& s) {   // String.trim
    size_t i = 0, j = s.size();
    while (i < j && (unsigned char)s[i] <= ' ') i++;
    while (j > i && (unsigned char)s[j - 1] <= ' ') j--;
    return s.substr(i, j - i);
}

// FileIOManager.loadScoringMatrix (FileIOManager.java:46-81), quirks included (rows in file order,
// dead header check, exactly 25 whitespace tokens per data line, > 24 rows is an error)
inline std::vector<int32_t> loadScoringMatrix(const std::string& path) {
    std::vector<int32_t> M(576, 0);
    int row = 0;
    for (const std::string& line : read_lines(path)) {
        if (!line.empty() && (line[0] == '#' || line[0] == ' ' || line[0] == '\t')) continue;
        std::vector<std::string> toks;
        if (line.empty()) toks.push_back("");
        else {
            if (java_ws((unsigned char)line[0])) toks.push_back("");
            size_t i = 0;
            while (i < line.size()) {
                while (i < line.size() && java_ws((unsigned char)line[i])) i++;
                if (i >= line.size()) break;
                size_t j = i;
                while (j < line.size() && !java_ws((unsigned char)line[j])) j++;
                toks.push_back(line.substr(i, j - i));
                i = j;
            }
        }
        if (toks.size() != 25)
            throw FileFormatException("Error in scoring matrix file: " + path +
                                      ". Scoring matrix should always have 24 columns (plus 1 column describing AAs).");
        if (row >= 24)
            throw FileFormatException("Error in scoring matrix file: " + path +
                                      ". Scoring matrix should always have 24 rows (plus 1 column describing AAs).");
        for (int c = 1; c < 25; c++) M[row * 24 + c - 1] = parse_int(toks[c]);
        row++;
    }
    return M;
}

// FileIOManager.loadUniqueSequencesFromFasta (FileIOManager.java:159-216)
inline std::vector<UniqueSequence> loadUniqueSequencesFromFasta(const std::string& path) {
    std::vector<std::string> order;                                       // LinkedHashMap: first-occurrence order
    std::unordered_map<std::string, std::vector<std::pair<std::string, int32_t>>> map;
    std::string sequence, label;
    int32_t count = 0;
    bool have = false;
    auto flush = [&]() {
        auto it = map.find(sequence);
        if (it == map.end()) { order.push_back(sequence); map[sequence] = {{label, count}}; }
        else {
            bool found = false;
            for (auto& kv : it->second) if (kv.first == label) { kv.second = wrap_add(kv.second, count); found = true; }
            if (!found) it->second.push_back({label, count});
        }
    };
    for (const std::string& line : read_lines(path)) {
        if (!line.empty() && line[0] == '>') {
            if (!sequence.empty()) { flush(); sequence.clear(); }
            std::string h = java_trim(line).substr(1);
            std::vector<std::string> parts;
            size_t i = 0;
            for (;;) {
                size_t j = h.find('|', i);
                parts.push_back(h.substr(i, j == std::string::npos ? std::string::npos : j - i));
                if (j == std::string::npos) break;
                i = j + 1;
            }
            while (parts.size() > 1 && parts.back().empty()) parts.pop_back();   // String.split drops trailing empties
            if (parts.size() >= 2) {
                count = decode_int(java_trim(parts[1]));
                if (count < 1) throw FileFormatException("Error while loading input file. Fasta header defines sequence count lower than 1.");
            } else count = 1;
            label = parts.size() >= 3 ? parts[2] : "no_label";
            have = true;
        } else {
            if (!have) throw FileFormatException("Error. Incorrect fasta format. Maybe header or sequence line missing?");
            sequence += java_trim(line);
        }
    }
    if (!have) throw FileFormatException("Error. Incorrect fasta format. Maybe header or sequence line missing?");
    flush();
    std::vector<UniqueSequence> out;
    for (auto& s : order) out.emplace_back(s, map[s]);
    return out;
}

// FileIOManager.loadUniqueSequencesFromTable (FileIOManager.java:227-255)
inline std::vector<UniqueSequence> loadUniqueSequencesFromTable(const std::string& path) {
    auto split = [](const std::string& l) {
        std::vector<std::string> p;
        size_t i = 0;
        for (;;) {
            size_t j = l.find('\t', i);
            p.push_back(l.substr(i, j == std::string::npos ? std::string::npos : j - i));
            if (j == std::string::npos) break;
            i = j + 1;
        }
        while (p.size() > 1 && p.back().empty()) p.pop_back();
        return p;
    };
    auto lines = read_lines(path);
    if (lines.empty()) throw FileFormatException("empty table");
    auto header = split(lines[0]);
    std::vector<UniqueSequence> out;
    for (size_t li = 1; li < lines.size(); li++) {
        auto parts = split(lines[li]);
        std::vector<std::pair<std::string, int32_t>> lab;
        for (size_t i = 1; i < parts.size(); i++) {
            int32_t v = decode_int(parts[i]);
            if (v != 0) lab.push_back({header.at(i), v});
        }
        out.emplace_back(parts[0], lab);
    }
    return out;
}

// ---------------------------------------------------------------- labels, ordering, defaults
// Hammock.getSortedLabels (Hammock.java:1586-1605): TreeMap with ValueComparator
// (FileIOManager.java:1464-1480, never returns 0) filled from a HashMap: count descending, equal
// counts in REVERSE HashMap iteration order.
inline std::vector<std::string> getSortedLabels(const std::vector<UniqueSequence>& seqs) {
    std::vector<std::string> keys;
    std::unordered_map<std::string, int32_t> cnt;
    for (auto& s : seqs)
        for (auto& kv : s.labels) {
            if (!cnt.count(kv.first)) { keys.push_back(kv.first); cnt[kv.first] = 0; }
            cnt[kv.first] = wrap_add(cnt[kv.first], kv.second);
        }
    std::vector<std::string> sorted;   // in-order content of the TreeMap
    for (auto& k : java_hashmap_order(keys)) {
        size_t pos = 0;                // after everything strictly larger, before everything <= (compare: a >= b -> -1)
        while (pos < sorted.size() && cnt[sorted[pos]] > cnt[k]) pos++;
        sorted.insert(sorted.begin() + pos, k);
    }
    return sorted;
}

// UniqueSequence.sortSequences (UniqueSequence.java:176-203)
inline void sortSequences(std::vector<UniqueSequence>& v, const std::string& order, const std::vector<std::string>& labels,
                          int64_t seed) {
    auto size_alpha_desc = [](const UniqueSequence& a, const UniqueSequence& b) {   // reverseOrder(SizeAlphabetic)
        int32_t sa = a.size(), sb = b.size();
        if (sa != sb) return sa > sb;
        return a.sequence > b.sequence;
    };
    if (order == "size") std::stable_sort(v.begin(), v.end(), size_alpha_desc);
    else if (order == "alphabetic")
        std::stable_sort(v.begin(), v.end(), [](const UniqueSequence& a, const UniqueSequence& b) { return a.sequence > b.sequence; });
    else if (order == "random") {   // Collections.shuffle(list, new Random(seed))
        JavaRandom rnd(seed);
        for (size_t i = v.size(); i > 1; i--) std::swap(v[i - 1], v[(size_t)rnd.nextInt((int32_t)i)]);
    } else if (order == "input") {
    } else {
        if (std::find(labels.begin(), labels.end(), order) == labels.end())
            throw DataException("Incorrect sequence order defined. Use one of: size, alphabetic, random, input, or a label");
        std::stable_sort(v.begin(), v.end(), size_alpha_desc);
        std::stable_sort(v.begin(), v.end(), [&](const UniqueSequence& a, const UniqueSequence& b) { return a.count(order) > b.count(order); });
    }
}

inline int32_t java_round(double x) { return (int32_t)std::floor(x + 0.5); }
inline double getMeanSequenceLength(const std::vector<UniqueSequence>& s) {   // Hammock.java:1554-1563
    int64_t sum = 0;
    for (auto& q : s) sum += (int64_t)q.codes.size();
    return (double)sum / (double)s.size();
}
inline int32_t setGreedyThreshold(const std::vector<UniqueSequence>& s) { return java_round(getMeanSequenceLength(s) * 1.7); }   // :1409-1413
inline int32_t checkMaxShift(const std::vector<UniqueSequence>& s, int32_t maxShift) {   // :1421-1427
    int32_t mn = INT32_MAX;
    for (auto& q : s) mn = std::min<int32_t>(mn, (int32_t)q.codes.size());
    return std::min(maxShift, mn - 1);
}
inline int32_t getMaxShift(const std::vector<UniqueSequence>& s) { return checkMaxShift(s, java_round(getMeanSequenceLength(s) / 4)); }   // :1429-1434
inline int32_t initialClustersLimit(const std::vector<UniqueSequence>& s) { return java_round((double)s.size() * 0.025); }   // :398-401

// ---------------------------------------------------------------- the clusterer (SequenceClusterer seam)
// == new LimitedGreedySequenceClusterer(new ShiftedScorer(matrix, shiftPenalty, maxShift), threshold,
//    maxClusters).cluster(sequences)   (Hammock.java:402-409)
struct GpuGreedySequenceClusterer {
    std::vector<int32_t> matrix;
    int32_t shiftPenalty, maxShift, threshold, maxClusters;
    int device = 0;

    std::vector<Cluster> cluster(const std::vector<UniqueSequence>& seqs) const {
        const int32_t n = (int32_t)seqs.size();
        std::vector<int32_t> off(n + 1, 0), ab(n), cid(std::max(n, 1)), rank(std::max(n, 1)), order(std::max(n, 1));
        std::vector<uint8_t> res;
        for (int32_t i = 0; i < n; i++) {
            res.insert(res.end(), seqs[i].codes.begin(), seqs[i].codes.end());
            off[i + 1] = (int32_t)res.size();
            ab[i] = seqs[i].size();
        }
        if (res.empty()) res.push_back(0);
        hmk_greedy_in in{n, res.data(), off.data(), ab.data(), matrix.data(), threshold, maxShift, shiftPenalty, maxClusters};
        hmk_greedy_out out{cid.data(), rank.data(), order.data(), 0, 0, -1};
        char err[512] = {0};
        int rc = hmk_greedy_cluster(&in, &out, device, err, sizeof err);
        if (rc == HMK_STATUS_SHIFT_TOO_BIG) throw DataException(err);
        if (rc == HMK_STATUS_NULL_CLUSTER) throw NullPointerException(out.error_step);
        if (rc == HMK_STATUS_BAD_RESIDUE) throw FileFormatException(err);
        if (rc != HMK_STATUS_OK) throw CudaException(std::string("hammock_b200: ") + err);
        // rebuild List<Cluster>: members by rank, clusters in result order
        std::vector<int> start(n + 1, 0);
        for (int32_t i = 0; i < n; i++) start[cid[i] + 1]++;
        for (int32_t i = 0; i < n; i++) start[i + 1] += start[i];
        std::vector<int> byRank(n);
        for (int32_t i = 0; i < n; i++) byRank[start[cid[i]] + rank[i]] = i;
        std::vector<Cluster> result;
        result.reserve(out.n_result);
        for (int32_t k = 0; k < out.n_result; k++) {
            Cluster c;
            c.id = order[k];
            for (int m = start[c.id]; m < start[c.id + 1]; m++) {
                c.members.push_back(byRank[m]);
                c.sizeSum = wrap_add(c.sizeSum, ab[byRank[m]]);
            }
            result.push_back(std::move(c));
        }
        return result;
    }
};

// ---------------------------------------------------------------- result files (FileIOManager.java)
static const char SEP = '\t';   // Hammock.CSV_SEPARATOR (Hammock.java:35)

// Cluster.compareTo reversed (Cluster.java:197-204): size descending, then id descending; stable
inline std::vector<const Cluster*> clustersLargestFirst(const std::vector<Cluster>& clusters) {
    std::vector<const Cluster*> v;
    for (auto& c : clusters) v.push_back(&c);
    std::stable_sort(v.begin(), v.end(), [](const Cluster* a, const Cluster* b) {
        if (a->size() != b->size()) return a->size() > b->size();
        return a->id > b->id;
    });
    return v;
}

inline void writeRow(std::ostream& w, const std::string& clusterId, const UniqueSequence& s, const std::string& alignment,
                     const std::vector<std::string>& labels) {
    w << clusterId << SEP << s.sequence << SEP << alignment << SEP << s.size();
    for (auto& l : labels) w << SEP << s.count(l);
    w << '\n';
}

// saveClusterSequencesToCsv (FileIOManager.java:398-404, 594-638): one row per sequence, clusters largest
// first, inside a cluster abundance desc then string desc.  The `alignment` column is the sequence
// itself for one-member clusters (Cluster.getFastaString) and "NA" for multi-member clusters: the
// reference fills it from the Clustal-Omega MSA it builds next (Hammock.java:414-426), which is outside
// this path; `cluster` mode accepts NA and rebuilds the MSAs (FileIOManager.java:351-357).
inline void saveClusterSequencesToCsv(const std::vector<Cluster>& clusters, const std::vector<UniqueSequence>& seqs,
                                      const std::string& path, const std::vector<std::string>& labels) {
    std::ofstream w(path);
    w << "cluster_id" << SEP << "sequence" << SEP << "alignment" << SEP << "sum";
    for (auto& l : labels) w << SEP << l;
    w << '\n';
    for (const Cluster* c : clustersLargestFirst(clusters)) {
        std::vector<int> m = c->members;
        std::stable_sort(m.begin(), m.end(), [&](int a, int b) {
            int32_t sa = seqs[a].size(), sb = seqs[b].size();
            if (sa != sb) return sa > sb;
            return seqs[a].sequence > seqs[b].sequence;
        });
        for (int i : m) writeRow(w, std::to_string(c->id), seqs[i], m.size() == 1 ? seqs[i].sequence : "NA", labels);
    }
}

// saveClusterSequencesToCsvOrdered (FileIOManager.java:371-374): same rows in the given sequence order
inline void saveClusterSequencesToCsvOrdered(const std::vector<Cluster>& clusters, const std::vector<UniqueSequence>& seqs,
                                             const std::vector<int>& sequenceOrder, const std::string& path,
                                             const std::vector<std::string>& labels) {
    std::vector<const Cluster*> of(seqs.size(), nullptr);
    for (auto& c : clusters) for (int i : c.members) of[i] = &c;
    std::ofstream w(path);
    w << "cluster_id" << SEP << "sequence" << SEP << "alignment" << SEP << "sum";
    for (auto& l : labels) w << SEP << l;
    w << '\n';
    for (int i : sequenceOrder) {
        const Cluster* c = of[i];
        if (c) writeRow(w, std::to_string(c->id), seqs[i], c->members.size() == 1 ? seqs[i].sequence : "NA", labels);
        else writeRow(w, "NA", seqs[i], "NA", labels);
    }
}

// SaveClustersToCsv (FileIOManager.java:649-676): main_sequence = most abundant member, ties alphabetically FIRST
inline void SaveClustersToCsv(const std::vector<Cluster>& clusters, const std::vector<UniqueSequence>& seqs,
                              const std::string& path, const std::vector<std::string>& labels) {
    std::ofstream w(path);
    w << "cluster_id" << SEP << "main_sequence" << SEP << "sum";
    for (auto& l : labels) w << SEP << l;
    w << '\n';
    for (const Cluster* c : clustersLargestFirst(clusters)) {
        int best = c->members[0];
        for (int i : c->members) {   // first element of sort(reverseOrder(UniqueSequence.compareTo)); stable
            int32_t si = seqs[i].size(), sb = seqs[best].size();
            if (si > sb || (si == sb && seqs[i].sequence < seqs[best].sequence)) best = i;
        }
        w << c->id << SEP << seqs[best].sequence << SEP << c->size();
        for (auto& l : labels) {
            int32_t sum = 0;
            for (int i : c->members) sum = wrap_add(sum, seqs[i].count(l));
            w << SEP << sum;
        }
        w << '\n';
    }
}

// saveInputStatistics (FileIOManager.java:709-729); no newline after the last line
inline void saveInputStatistics(const std::vector<UniqueSequence>& seqs, const std::vector<std::string>& labels, const std::string& path) {
    std::ofstream w(path);
    for (auto& l : labels) w << SEP << l;
    w << '\n' << "total_count";
    for (auto& l : labels) { int32_t t = 0; for (auto& s : seqs) t = wrap_add(t, s.count(l)); w << SEP << t; }
    w << '\n' << "unique_count";
    for (auto& l : labels) { int32_t u = 0; for (auto& s : seqs) u += s.has(l) ? 1 : 0; w << SEP << u; }
}

}  // namespace hammock
