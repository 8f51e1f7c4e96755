/*
 * GpuGreedySequenceClusterer -- drop-in SequenceClusterer that runs Hammock's greedy initial
 * clustering on a B200 through libhammock_b200.so (C ABI: include/hammock_b200.h).
 *
 * SOURCE ONLY: the build image has no JDK, so this file has never been compiled; it documents
 * the binding a Hammock maintainer would add (see INTEGRATION.md).  It needs Java 22+ (the
 * java.lang.foreign API); a JNI variant is sketched in INTEGRATION.md for Java 7-21.
 *
 * It replaces, at Hammock.java:402-403,
 *     AligningSequenceScorer scorer = new ShiftedScorer(scoringMatrix, shiftPenalty, maxShift);
 *     SequenceClusterer clusterer   = new LimitedGreedySequenceClusterer(scorer, threshold, limit);
 * by
 *     SequenceClusterer clusterer = new GpuGreedySequenceClusterer(scoringMatrix, shiftPenalty,
 *                                        maxShift, sequenceClusteringThreshold, initialClustersLimit, 0);
 * Everything before (parsing, sortSequences) and after (Clustal, writers) is unchanged.
 */
package cz.krejciadam.hammock;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.util.ArrayList;
import java.util.List;
import java.util.concurrent.ExecutionException;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public class GpuGreedySequenceClusterer implements SequenceClusterer {

    /* struct hmk_greedy_in (include/hammock_b200.h) */
    private static final StructLayout IN = MemoryLayout.structLayout(
            JAVA_INT.withName("n"), MemoryLayout.paddingLayout(4),
            ADDRESS.withName("residues"), ADDRESS.withName("offsets"), ADDRESS.withName("abundance"),
            ADDRESS.withName("matrix"),
            JAVA_INT.withName("threshold"), JAVA_INT.withName("max_shift"),
            JAVA_INT.withName("shift_penalty"), JAVA_INT.withName("max_clusters"));
    /* struct hmk_greedy_out */
    private static final StructLayout OUT = MemoryLayout.structLayout(
            ADDRESS.withName("cluster_id"), ADDRESS.withName("member_rank"), ADDRESS.withName("result_order"),
            JAVA_INT.withName("n_result"), JAVA_INT.withName("n_multi"), JAVA_INT.withName("error_step"),
            MemoryLayout.paddingLayout(4));

    private static final MethodHandle HMK_GREEDY_CLUSTER, HMK_GREEDY_CLUSTER_MULTI;

    static {
        System.loadLibrary("hammock_b200");          // libhammock_b200.so on java.library.path
        HMK_GREEDY_CLUSTER = Linker.nativeLinker().downcallHandle(
                SymbolLookup.loaderLookup().find("hmk_greedy_cluster").orElseThrow(),
                FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG));
        /* int hmk_greedy_cluster_multi(in, out, const int32_t* devices, int32_t n_gpus, errbuf, errlen): the same call
         * on several GPUs of this JVM -- worker threads and the NCCL communicator live inside the library */
        HMK_GREEDY_CLUSTER_MULTI = Linker.nativeLinker().downcallHandle(
                SymbolLookup.loaderLookup().find("hmk_greedy_cluster_multi").orElseThrow(),
                FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG));
    }

    private final int[][] scoringMatrix;
    private final int shiftPenalty, maxShift, threshold, maxClusters;
    private final int[] devices;

    public GpuGreedySequenceClusterer(int[][] scoringMatrix, int shiftPenalty, int maxShift,
                                      int threshold, int maxClusters, int device) {
        this(scoringMatrix, shiftPenalty, maxShift, threshold, maxClusters, new int[]{device});
    }

    /** several GPUs of this machine, e.g. {0, 1, 2, 3, 4, 5, 6, 7}: identical result, the partner search is sharded */
    public GpuGreedySequenceClusterer(int[][] scoringMatrix, int shiftPenalty, int maxShift,
                                      int threshold, int maxClusters, int[] devices) {
        this.scoringMatrix = scoringMatrix;
        this.shiftPenalty = shiftPenalty;
        this.maxShift = maxShift;
        this.threshold = threshold;
        this.maxClusters = maxClusters;
        this.devices = devices.clone();
    }

    @Override
    public List<Cluster> cluster(List<UniqueSequence> sequences)
            throws InterruptedException, ExecutionException, DataException {
        final int n = sequences.size();
        try (Arena arena = Arena.ofConfined()) {
            int total = 0;
            for (UniqueSequence s : sequences) total += s.getSequence().length;
            MemorySegment residues = arena.allocate(Math.max(total, 1));
            MemorySegment offsets = arena.allocate(JAVA_INT, n + 1L);
            MemorySegment abundance = arena.allocate(JAVA_INT, Math.max(n, 1));
            int pos = 0;
            for (int i = 0; i < n; i++) {
                offsets.setAtIndex(JAVA_INT, i, pos);
                for (int r : sequences.get(i).getSequence()) residues.set(JAVA_BYTE, pos++, (byte) r);
                abundance.setAtIndex(JAVA_INT, i, sequences.get(i).size());
            }
            offsets.setAtIndex(JAVA_INT, n, pos);
            MemorySegment matrix = arena.allocate(JAVA_INT, 24 * 24);
            for (int a = 0; a < 24; a++)
                for (int b = 0; b < 24; b++) matrix.setAtIndex(JAVA_INT, a * 24L + b, scoringMatrix[a][b]);

            MemorySegment in = arena.allocate(IN);
            in.set(JAVA_INT, 0, n);
            in.set(ADDRESS, 8, residues);
            in.set(ADDRESS, 16, offsets);
            in.set(ADDRESS, 24, abundance);
            in.set(ADDRESS, 32, matrix);
            in.set(JAVA_INT, 40, threshold);
            in.set(JAVA_INT, 44, maxShift);
            in.set(JAVA_INT, 48, shiftPenalty);
            in.set(JAVA_INT, 52, maxClusters);

            MemorySegment clusterId = arena.allocate(JAVA_INT, Math.max(n, 1));
            MemorySegment memberRank = arena.allocate(JAVA_INT, Math.max(n, 1));
            MemorySegment resultOrder = arena.allocate(JAVA_INT, Math.max(n, 1));
            MemorySegment out = arena.allocate(OUT);
            out.set(ADDRESS, 0, clusterId);
            out.set(ADDRESS, 8, memberRank);
            out.set(ADDRESS, 16, resultOrder);
            MemorySegment err = arena.allocate(512);

            int rc;
            try {
                if (devices.length == 1) {
                    rc = (int) HMK_GREEDY_CLUSTER.invokeExact(in, out, devices[0], err, 512L);
                } else {
                    MemorySegment devs = arena.allocate(JAVA_INT, devices.length);
                    for (int i = 0; i < devices.length; i++) devs.setAtIndex(JAVA_INT, i, devices[i]);
                    rc = (int) HMK_GREEDY_CLUSTER_MULTI.invokeExact(in, out, devs, devices.length, err, 512L);
                }
            } catch (Throwable t) {
                throw new ExecutionException(t);
            }
            switch (rc) {
                case 0: break;
                case 1: throw new DataException(err.getString(0));              // ShiftedScorer.java:59-62
                case 2: throw new NullPointerException(                          // LimitedGreedy...java:104,108
                        "greedy phase 1, step " + out.get(JAVA_INT, 32));
                case 3: throw new DataException("invalid residue code");          // cannot happen: UniqueSequence validated
                default: throw new ExecutionException(new RuntimeException("hammock_b200: " + err.getString(0)));
            }

            /* rebuild List<Cluster>: members in insertion order, clusters in result_order */
            final int nResult = out.get(JAVA_INT, 24);
            int[] start = new int[n + 1];                     // members per cluster id (counting sort by id)
            for (int i = 0; i < n; i++) start[clusterId.getAtIndex(JAVA_INT, i) + 1]++;
            for (int i = 0; i < n; i++) start[i + 1] += start[i];
            int[] byRank = new int[n];
            for (int i = 0; i < n; i++)
                byRank[start[clusterId.getAtIndex(JAVA_INT, i)] + memberRank.getAtIndex(JAVA_INT, i)] = i;
            List<Cluster> result = new ArrayList<>(nResult);
            for (int k = 0; k < nResult; k++) {
                int id = resultOrder.getAtIndex(JAVA_INT, k);
                List<UniqueSequence> first = new ArrayList<>();
                first.add(sequences.get(byRank[start[id]]));
                Cluster cl = new Cluster(first, id);                              // Cluster.java:31-41
                for (int m = start[id] + 1; m < start[id + 1]; m++)
                    cl.insert(sequences.get(byRank[m]));                          // Cluster.java:50-63
                result.add(cl);
            }
            return result;
        }
    }
}
