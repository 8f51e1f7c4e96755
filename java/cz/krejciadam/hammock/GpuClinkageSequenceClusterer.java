/*
 * GpuClinkageSequenceClusterer -- drop-in SequenceClusterer for Hammock's EXACT complete-linkage initial clustering
 * (the default initial stage for up to 10 000 unique sequences, Hammock.java:371-373) on a B200 through
 * libhammock_b200.so: int hmk_clinkage_cluster(in, out, device, errbuf, errlen)   (include/hammock_b200.h).
 *
 * SOURCE ONLY (no JDK in the build image; never compiled).  Java 22+ (java.lang.foreign).
 *
 * It replaces, at Hammock.java:458-459,
 *     ShiftedScorer scorer = new ShiftedScorer(scoringMatrix, shiftPenalty, maxShift);
 *     clusterer = new ClinkageSequenceClusterer(scorer, sequenceClusteringThreshold);
 * by
 *     clusterer = new GpuClinkageSequenceClusterer(scoringMatrix, shiftPenalty, maxShift, sequenceClusteringThreshold, 0);
 *
 * The library returns Cluster.getId() per sequence (i + 1 for singletons, n + 2, n + 3, ... for merged clusters, as
 * ClinkageSequenceClusterer.java:48-54,96 numbers them), the position of every sequence in getSequences(), and the ids
 * of the returned list in the order new ArrayList(readyClusters) has on an OpenJDK 8+ runtime (:119-123).
 */
package cz.krejciadam.hammock;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;
import java.util.concurrent.ExecutionException;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public class GpuClinkageSequenceClusterer implements SequenceClusterer {

    private static final MethodHandle HMK_CLINKAGE_CLUSTER;

    static {
        System.loadLibrary("hammock_b200");
        HMK_CLINKAGE_CLUSTER = Linker.nativeLinker().downcallHandle(
                SymbolLookup.loaderLookup().find("hmk_clinkage_cluster").orElseThrow(),
                FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG));
    }

    private final int[][] scoringMatrix;
    private final int shiftPenalty, maxShift, threshold, device;

    public GpuClinkageSequenceClusterer(int[][] scoringMatrix, int shiftPenalty, int maxShift, int threshold, int device) {
        this.scoringMatrix = scoringMatrix;
        this.shiftPenalty = shiftPenalty;
        this.maxShift = maxShift;
        this.threshold = threshold;
        this.device = device;
    }

    @Override
    public List<Cluster> cluster(List<UniqueSequence> sequences) throws InterruptedException, ExecutionException, DataException {
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
            MemorySegment in = arena.allocate(56);           /* struct hmk_greedy_in, see GpuGreedySequenceClusterer */
            in.set(JAVA_INT, 0, n);
            in.set(ADDRESS, 8, residues);
            in.set(ADDRESS, 16, offsets);
            in.set(ADDRESS, 24, abundance);
            in.set(ADDRESS, 32, matrix);
            in.set(JAVA_INT, 40, threshold);
            in.set(JAVA_INT, 44, maxShift);
            in.set(JAVA_INT, 48, shiftPenalty);
            in.set(JAVA_INT, 52, 0);                         /* max_clusters: unused by this clusterer */
            MemorySegment clusterId = arena.allocate(JAVA_INT, Math.max(n, 1));
            MemorySegment memberRank = arena.allocate(JAVA_INT, Math.max(n, 1));
            MemorySegment resultOrder = arena.allocate(JAVA_INT, Math.max(n, 1));
            MemorySegment out = arena.allocate(40);          /* struct hmk_greedy_out */
            out.set(ADDRESS, 0, clusterId);
            out.set(ADDRESS, 8, memberRank);
            out.set(ADDRESS, 16, resultOrder);
            MemorySegment err = arena.allocate(512);
            int rc;
            try {
                rc = (int) HMK_CLINKAGE_CLUSTER.invokeExact(in, out, device, err, 512L);
            } catch (Throwable t) {
                throw new ExecutionException(t);
            }
            switch (rc) {
                case 0: break;
                case 1: throw new DataException(err.getString(0));                       // ShiftedScorer.java:59-62
                case 5: throw new java.util.NoSuchElementException(err.getString(0));    // empty input, :116
                case 6: throw new ExecutionException(new UnsupportedOperationException(err.getString(0)));
                default: throw new ExecutionException(new RuntimeException("hammock_b200: " + err.getString(0)));
            }
            /* members of every cluster in getSequences() order */
            Map<Integer, UniqueSequence[]> members = new HashMap<>();
            int[] count = new int[2 * n + 3];
            for (int i = 0; i < n; i++) count[clusterId.getAtIndex(JAVA_INT, i)]++;
            for (int i = 0; i < n; i++) {
                int id = clusterId.getAtIndex(JAVA_INT, i);
                members.computeIfAbsent(id, k -> new UniqueSequence[count[k]])[memberRank.getAtIndex(JAVA_INT, i)] = sequences.get(i);
            }
            final int nResult = out.get(JAVA_INT, 24);
            List<Cluster> result = new ArrayList<>(nResult);
            for (int k = 0; k < nResult; k++) {
                int id = resultOrder.getAtIndex(JAVA_INT, k);
                List<UniqueSequence> seqs = new ArrayList<>(java.util.Arrays.asList(members.get(id)));
                result.add(new Cluster(seqs, id));                                       // Cluster.java:31-41
            }
            return result;
        }
    }
}
