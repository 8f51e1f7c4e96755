// Test harness (CPU tests only): the C++ host side either side of the GPU call, without the GPU call.
// input fasta -> loader, labels, sortSequences -> [a clustering read from a file, computed by the ORACLE in the test]
// -> rebuildClusters -> the four result files.  The driver (hammock_greedy.cpp) runs exactly these functions around
// hmk_greedy_cluster; this harness never links libhammock_b200.
//   usage: writers_harness <in.fa> <order> <seed> <result.txt> <outdir/>
//   result.txt: n n_result, then n cluster ids, n ranks, n_result ids of the result list (clustering order)
#include <chrono>
#include <cstdlib>
#include <iostream>

#include "../../hammock_b200/host_cpp/hammock_host.hpp"

using namespace hammock;

int main(int argc, char** argv) {
    if (argc != 6) return 2;
    auto tick = std::chrono::steady_clock::now();
    const bool timing = std::getenv("HARNESS_TIMES") != nullptr;     // stage times on stderr, no sequence dump
    auto lap = [&](const char* what) {
        auto now = std::chrono::steady_clock::now();
        if (timing) std::cerr << "host time " << what << ": " << std::chrono::duration<double, std::milli>(now - tick).count() << " ms\n";
        tick = now;
    };
    try {
        std::vector<UniqueSequence> sequences = loadUniqueSequencesFromFasta(argv[1]);
        std::vector<std::string> labels = getSortedLabels(sequences);
        std::vector<int> cameFrom;
        sortSequences(sequences, argv[2], labels, decode_int(argv[3]), &cameFrom);
        lap("load + labels + sort");
        std::ifstream f(argv[4]);
        int32_t n = 0, nResult = 0;
        f >> n >> nResult;
        if (n != (int32_t)sequences.size()) { std::cerr << "size mismatch\n"; return 3; }
        std::vector<int32_t> cid(n), rank(n), order(nResult), ab(n);
        for (auto& v : cid) f >> v;
        for (auto& v : rank) f >> v;
        for (auto& v : order) f >> v;
        for (int32_t i = 0; i < n; i++) ab[i] = sequences[i].size();
        {   // the arrays the clusterers hand to the C ABI
            const PackedSequences packed(sequences);
            std::ofstream pf(std::string(argv[5]) + "packed.txt");
            for (auto v : packed.off) pf << v << ' ';
            pf << '\n';
            for (auto v : packed.ab) pf << v << ' ';
            pf << '\n';
            for (auto v : packed.res) pf << (int)v << ' ';
            pf << '\n';
        }
        lap("read the clustering (test input)");
        std::vector<Cluster> clusters = rebuildClusters(n, cid.data(), rank.data(), order.data(), nResult, ab.data());
        std::vector<int> inputOrder(sequences.size());
        for (size_t i = 0; i < cameFrom.size(); i++) inputOrder[cameFrom[i]] = (int)i;
        lap("rebuildClusters");
        const std::string d = argv[5];
        saveInputStatistics(sequences, labels, d + "input_statistics.tsv");
        lap("saveInputStatistics");
        saveClusterSequencesToCsv(clusters, sequences, d + "initial_clusters_sequences.tsv", labels);
        lap("saveClusterSequencesToCsv");
        saveClusterSequencesToCsvOrdered(clusters, sequences, inputOrder, d + "initial_clusters_sequences_original_order.tsv", labels);
        lap("saveClusterSequencesToCsvOrdered");
        SaveClustersToCsv(clusters, sequences, d + "initial_clusters.tsv", labels);
        lap("SaveClustersToCsv");
        if (!timing)
            for (auto& s : sequences) std::cout << s.sequence << '\t' << s.size() << '\n';
        return 0;
    } catch (const HammockException& e) {
        std::cerr << "Error: " << e.what() << "\n";
        return 1;
    }
}
