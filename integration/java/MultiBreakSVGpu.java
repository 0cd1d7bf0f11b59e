/*
 * MultiBreakSVGpu — drop-in subclass of the reference's MultiBreakSV whose runMCMC() (MultiBreakSV.java:773-877) runs on a
 * B200 through libmbsv_b200.so, bound with the Panama Foreign Function & Memory API (java.lang.foreign, JDK 22+).
 *
 * NOT COMPILED IN THIS REPOSITORY: neither the build image nor the GPU boxes have a JDK.  It is the binding a maintainer
 * adds next to the reference's sources (default package, like them):
 *     javac --release 22 -cp bin/MultiBreakSV.jar MultiBreakSVGpu.java
 *     java --enable-native-access=ALL-UNNAMED -Djava.library.path=<dir of libmbsv_b200.so> -cp .:bin/MultiBreakSV.jar \
 *          MultiBreakSVGpu --clusterfile ... --assignmentfile ... --experimentfile ... --numiterations N --perr P
 * Struct layouts follow include/mbsv_b200.h; tests/test_abi_cpu.py pins their sizeof/offsetof.
 */
import java.io.BufferedWriter;
import java.io.FileWriter;
import java.io.IOException;
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

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public class MultiBreakSVGpu extends MultiBreakSV {

    static final Linker LINKER = Linker.nativeLinker();
    static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.mapLibraryName("mbsv_b200"), Arena.global());

    static MethodHandle h(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
    }

    static final MethodHandle ABI_VERSION = h("mbsv_abi_version", FunctionDescriptor.of(JAVA_INT));
    static final MethodHandle CREATE = h("mbsv_problem_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle RUN = h("mbsv_run", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle DESTROY = h("mbsv_problem_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    // the library keeps large device blocks of destroyed problems for the next one; TRIM gives them back (device < 0: all)
    static final MethodHandle TRIM = h("mbsv_trim", FunctionDescriptor.of(JAVA_INT, JAVA_INT));
    static final MethodHandle LAST_ERROR = h("mbsv_last_error", FunctionDescriptor.of(ADDRESS));
    // giant component on every GPU of the box: chains split over the devices, tallies pooled with one ncclReduce
    // (include/mbsv_b200.h: mbsv_run_multi; problems = one handle per device, created from the same descriptor)
    static final MethodHandle RUN_MULTI = h("mbsv_run_multi",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle DEVICE_COUNT = h("mbsv_device_count", FunctionDescriptor.of(JAVA_INT));

    public int device = 0;
    public int mode = 0;   // MBSV_MODE_REPLAY_EXACT

    public MultiBreakSVGpu(String[] args) { super(args); }

    static String lastError() throws Throwable {
        MemorySegment p = (MemorySegment) LAST_ERROR.invokeExact();
        return p.reinterpret(4096).getString(0);
    }

    static void check(int rc) throws Throwable {
        if (rc == -5) throw new Error(lastError());           // the reference's own Error conditions (:1122-1130,1142-1144)
        if (rc != 0) throw new IllegalStateException("libmbsv_b200: status " + rc + ": " + lastError());
    }

    static MemorySegment ints(Arena a, List<Integer> v) {
        MemorySegment s = a.allocate(JAVA_INT, Math.max(1, v.size()));
        for (int i = 0; i < v.size(); i++) s.setAtIndex(JAVA_INT, i, v.get(i));
        return s;
    }

    /** The reference's object graph flattened as include/mbsv_b200.h describes (one subproblem). */
    final class Packed {
        final List<String> fragIds = new ArrayList<>(fragments.keySet());          // iteration order matters (:999,1401)
        final List<String> clusterIds = new ArrayList<>(breakpoints.keySet());     // (:750,892)
        final List<String> expLabels = new ArrayList<>(experiments.keySet());
        final List<Integer> fragAlnPtr = new ArrayList<>(), alnNumerrors = new ArrayList<>(), alnLen = new ArrayList<>(),
                alnExp = new ArrayList<>(), alnEspPtr = new ArrayList<>(), alnEspCluster = new ArrayList<>(),
                alnEspLeaf = new ArrayList<>(), supportsOrder = new ArrayList<>(), histPtr = new ArrayList<>(),
                leafPtr = new ArrayList<>(), leafOwner = new ArrayList<>(), rootPtr = new ArrayList<>(),
                rootLeafPtr = new ArrayList<>(), rootLeaf = new ArrayList<>(), svCluster = new ArrayList<>(),
                svFragPtr = new ArrayList<>(), svFrag = new ArrayList<>(), svNumchanged = new ArrayList<>();
        final byte[] clusterKind;
        MemorySegment desc;

        Packed(Arena arena) {
            Map<String, Integer> fragIdx = new HashMap<>(), clIdx = new HashMap<>(), expIdx = new HashMap<>();
            for (int i = 0; i < fragIds.size(); i++) fragIdx.put(fragIds.get(i), i);
            for (int i = 0; i < clusterIds.size(); i++) clIdx.put(clusterIds.get(i), i);
            for (int i = 0; i < expLabels.size(); i++) expIdx.put(expLabels.get(i), i);
            int C = clusterIds.size();
            clusterKind = new byte[C];
            // diagrams: leaves, owners, roots
            List<Map<String, Integer>> leafSlot = new ArrayList<>();
            leafPtr.add(0); rootPtr.add(0); rootLeafPtr.add(0); histPtr.add(0);
            for (int c = 0; c < C; c++) {
                MultiBreakSVDiagram d = diagrams.get(clusterIds.get(c));
                clusterKind[c] = (byte) (d.isCluster ? 0 : 1);
                Map<String, Integer> slots = new HashMap<>();
                for (String esp : d.leaves.keySet()) {
                    slots.put(esp, leafOwner.size() - leafPtr.get(c));
                    String owner = esp2fragmentID.get(esp);
                    leafOwner.add(owner != null && fragIdx.containsKey(owner) ? fragIdx.get(owner) : -1);
                }
                leafSlot.add(slots);
                leafPtr.add(leafOwner.size());
                for (DiagramNode root : d.roots) {
                    for (String esp : root.esps) rootLeaf.add(slots.get(esp));
                    rootLeafPtr.add(rootLeaf.size());
                }
                rootPtr.add(rootLeafPtr.size() - 1);
                histPtr.add(histPtr.get(c) + maxSupport + 1);       // uniform stride: supportslong has maxSupport+1 bins
            }
            // A.supports.keySet() order: a fresh matrix is filled by iterating breakpoints (AM.java:53-66)
            for (String cid : new MultiBreakSVAssignmentMatrix().supports.keySet()) supportsOrder.add(clIdx.get(cid));
            // alignments
            fragAlnPtr.add(0); alnEspPtr.add(0);
            for (int f = 0; f < fragIds.size(); f++) {
                for (FragmentAlignment fa : fragments.get(fragIds.get(f)).fragmentalignments) {
                    alnNumerrors.add(fa.numerrors); alnLen.add(fa.len); alnExp.add(expIdx.get(fa.exp.label));
                    for (String esp : fa.alignments) {
                        int c = clIdx.get(esp2cid.get(esp));
                        Integer slot = leafSlot.get(c).get(esp);
                        boolean counts = slot != null && leafOwner.get(leafPtr.get(c) + slot) == f;
                        alnEspCluster.add(counts ? c : -1);
                        alnEspLeaf.add(counts ? slot : -1);
                    }
                    alnEspPtr.add(alnEspCluster.size());
                }
                fragAlnPtr.add(alnNumerrors.size());
            }
            // AlterSV lists (:748-771, :1303-1330)
            svFragPtr.add(0);
            for (String cid : bpsWithNonSingletonSupport) {
                int c = clIdx.get(cid), numchanged = 0;
                List<Integer> uniq = new ArrayList<>();
                for (int l = leafPtr.get(c); l < leafPtr.get(c + 1); l++) {
                    int f = leafOwner.get(l);
                    if (f >= 0 && fragAlnPtr.get(f + 1) - fragAlnPtr.get(f) == 1) {
                        numchanged++;
                        if (!uniq.contains(f)) uniq.add(f);
                    }
                }
                svCluster.add(c); svNumchanged.add(numchanged); svFrag.addAll(uniq); svFragPtr.add(svFrag.size());
            }
            // mbsv_problem_desc: 14 int32 scalars, then 25 pointers (include/mbsv_b200.h)
            int E = expLabels.size();
            desc = arena.allocate(14 * 4 + 25 * 8, 8);
            int[] scalars = {1, 1, fragIds.size(), alnNumerrors.size(), alnEspCluster.size(), C, E, maxSupport,
                    svCluster.size(), svFrag.size(), leafOwner.size(), rootLeafPtr.size() - 1, rootLeaf.size(), histPtr.get(C)};
            for (int i = 0; i < scalars.length; i++) desc.set(JAVA_INT, 4L * i, scalars[i]);
            MemorySegment pseq = arena.allocate(JAVA_DOUBLE, E), lam = arena.allocate(JAVA_DOUBLE, E),
                    cov = arena.allocate(JAVA_DOUBLE, (long) E * (maxSupport + 1)), kind = arena.allocate(JAVA_BYTE, Math.max(1, C));
            for (int x = 0; x < E; x++) {
                Experiment ex = experiments.get(expLabels.get(x));
                pseq.setAtIndex(JAVA_DOUBLE, x, ex.pseq); lam.setAtIndex(JAVA_DOUBLE, x, ex.lambdaD);
                for (int k = 0; k <= maxSupport; k++) cov.setAtIndex(JAVA_DOUBLE, (long) x * (maxSupport + 1) + k, ex.logprobcov[k]);
            }
            for (int c = 0; c < C; c++) kind.set(JAVA_BYTE, c, clusterKind[c]);
            MemorySegment[] ptrs = {
                ints(arena, List.of(0, fragIds.size())), ints(arena, List.of(0, C)), ints(arena, List.of(0, svCluster.size())),
                ints(arena, fragAlnPtr), ints(arena, alnNumerrors), ints(arena, alnLen), ints(arena, alnExp), ints(arena, alnEspPtr),
                ints(arena, alnEspCluster), ints(arena, alnEspLeaf), kind, ints(arena, supportsOrder), ints(arena, histPtr),
                ints(arena, leafPtr), ints(arena, leafOwner), ints(arena, rootPtr), ints(arena, rootLeafPtr), ints(arena, rootLeaf),
                ints(arena, svCluster), ints(arena, svFragPtr), ints(arena, svFrag), ints(arena, svNumchanged), pseq, lam, cov};
            for (int i = 0; i < ptrs.length; i++) desc.set(ADDRESS, 56L + 8L * i, ptrs[i]);
        }
    }

    @Override
    public MultiBreakSVAssignmentMatrix runMCMC(String file) throws IOException, ClassNotFoundException {
        try (Arena arena = Arena.ofConfined()) {
            if ((int) ABI_VERSION.invokeExact() != 1) throw new IllegalStateException("libmbsv_b200 ABI mismatch");
            Packed p = new Packed(arena);
            int F = p.fragIds.size(), A = p.alnNumerrors.size(), H = p.histPtr.get(p.clusterIds.size());
            MemorySegment out = arena.allocate(ADDRESS);
            check((int) CREATE.invokeExact(p.desc, device, out));
            MemorySegment prob = out.get(ADDRESS, 0);
            boolean trace = file != null;
            // mbsv_run_params {int mode,n_chains,num_iters,burnin; double perr; long seed; int device,java_int_wrap;
            //                  const int* init_assign; int trace,memo_states}
            MemorySegment rp = arena.allocate(56, 8);
            rp.set(JAVA_INT, 0, mode); rp.set(JAVA_INT, 4, 1); rp.set(JAVA_INT, 8, numiters); rp.set(JAVA_INT, 12, burnin);
            rp.set(JAVA_DOUBLE, 16, perr); rp.set(JAVA_LONG, 24, 123456L); rp.set(JAVA_INT, 32, device); rp.set(JAVA_INT, 36, 1);
            rp.set(ADDRESS, 40, MemorySegment.NULL); rp.set(JAVA_INT, 48, trace ? 1 : 0); rp.set(JAVA_INT, 52, 65536);
            // mbsv_results: 14 pointers then double kernel_ms
            MemorySegment alnCount = arena.allocate(JAVA_INT, Math.max(1, A)), fragErr = arena.allocate(JAVA_INT, F),
                    hist = arena.allocate(JAVA_INT, Math.max(1, H)), fin = arena.allocate(JAVA_INT, F),
                    flp = arena.allocate(JAVA_DOUBLE, 1), counters = arena.allocate(JAVA_LONG, 5),
                    trLp = trace ? arena.allocate(JAVA_DOUBLE, numiters) : MemorySegment.NULL,
                    trF = trace ? arena.allocate(JAVA_INT, numiters) : MemorySegment.NULL,
                    trA = trace ? arena.allocate(JAVA_INT, numiters) : MemorySegment.NULL,
                    init = arena.allocate(JAVA_INT, F);
            MemorySegment rs = arena.allocate(14 * 8 + 8, 8);
            MemorySegment[] r = {alnCount, fragErr, hist, fin, flp, counters, MemorySegment.NULL, MemorySegment.NULL, trLp, trF, trA,
                    init, MemorySegment.NULL, MemorySegment.NULL};
            for (int i = 0; i < r.length; i++) rs.set(ADDRESS, 8L * i, r[i]);
            int rc = (int) RUN.invokeExact(prob, rp, rs);
            int rcd = (int) DESTROY.invokeExact(prob);
            check(rc);
            check(rcd);
            // ---- what :846-870 leaves behind in the reference's objects ----
            for (int f = 0; f < F; f++) {
                Fragment fr = fragments.get(p.fragIds.get(f));
                fr.avgNoassignmentlong = Integer.toUnsignedLong(fragErr.getAtIndex(JAVA_INT, f)) / (double) burnin;
                for (int j = 0; j < fr.fragmentalignments.size(); j++)
                    fr.avgAssignmentslong.set(j, Integer.toUnsignedLong(alnCount.getAtIndex(JAVA_INT, p.fragAlnPtr.get(f) + j)) / (double) burnin);
            }
            for (int c = 0; c < p.clusterIds.size(); c++) {
                double[] s = breakpoints.get(p.clusterIds.get(c)).supportslong;
                for (int k = 1; k < s.length; k++) s[k] = Integer.toUnsignedLong(hist.getAtIndex(JAVA_INT, p.histPtr.get(c) + k));
            }
            proposalCounts[0] = (int) counters.getAtIndex(JAVA_LONG, 0); moveCounts[0] = (int) counters.getAtIndex(JAVA_LONG, 1);
            proposalCounts[1] = (int) counters.getAtIndex(JAVA_LONG, 2); moveCounts[1] = (int) counters.getAtIndex(JAVA_LONG, 3);
            MultiBreakSVAssignmentMatrix A_ = new MultiBreakSVAssignmentMatrix();
            for (int f = 0; f < F; f++) A_.assignments.put(p.fragIds.get(f), fin.getAtIndex(JAVA_INT, f));
            A_.setBreakpoints(diagrams);
            A_.logprob = flp.getAtIndex(JAVA_DOUBLE, 0);
            if (trace) writeTrace(file, p, A_, init, trLp, trF, trA);
            return A_;
        } catch (IOException | Error e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** mcmc.txt (:779-828) and the counts/logprobs maps writeStateCounts (:1934-1949) prints, from the accepted-move log. */
    void writeTrace(String file, Packed p, MultiBreakSVAssignmentMatrix A_, MemorySegment init, MemorySegment trLp,
                    MemorySegment trF, MemorySegment trA) throws IOException {
        int numbps = p.clusterIds.size() * p.expLabels.size();    // sum of supports.get(cID).size(), a constant (:811-815)
        int[] cur = new int[p.fragIds.size()];
        for (int f = 0; f < cur.length; f++) cur[f] = init.getAtIndex(JAVA_INT, f);
        Map<String, Integer> pos = new HashMap<>();
        for (int f = 0; f < cur.length; f++) pos.put(p.fragIds.get(f), f);
        try (BufferedWriter w = new BufferedWriter(new FileWriter(file))) {
            w.write("Iteration\tLogProb\tNumBreakpoints\tMadeMove?\n");
            for (int i = 0; i < numiters; i++) {
                int mf = trF.getAtIndex(JAVA_INT, i);
                if (mf >= 0) cur[mf] = trA.getAtIndex(JAVA_INT, i);
                else if (mf <= -2) {
                    int k = -mf - 2;
                    for (int q = p.svFragPtr.get(k); q < p.svFragPtr.get(k + 1); q++) cur[p.svFrag.get(q)] = cur[p.svFrag.get(q)] == -1 ? 0 : -1;
                }
                double lp = trLp.getAtIndex(JAVA_DOUBLE, i);
                StringBuilder key = new StringBuilder().append(cur[pos.get(A_.keys.get(0))]);
                for (int k = 1; k < A_.keys.size(); k++) key.append('\t').append(cur[pos.get(A_.keys.get(k))]);
                String s = key.toString();
                counts.merge(s, 1, Integer::sum);
                logprobs.putIfAbsent(s, lp);
                w.write(i + "\t" + lp + "\t" + (i % skip == 0 ? numbps : 0) + "\t" + (mf != -1 ? 1 : 0) + "\n");
            }
        }
    }

    /** Clone of MultiBreakSV.main (:114-153) with the subclass. */
    public static void main(String[] args) throws Exception {
        long starttime = System.nanoTime();
        MultiBreakSVGpu c2g = new MultiBreakSVGpu(args);
        c2g.getFragmentAlignments(false);
        c2g.constructClusterDiagrams(false);
        c2g.precomputePriors();
        c2g.fillNonSingletonArray();
        MultiBreakSVAssignmentMatrix A = c2g.runMCMC(SAVE_SPACE ? null : c2g.outputname + ".mcmc.txt");
        c2g.writeFinalFiles();
        if (!SAVE_SPACE) c2g.writeStateCounts(A.keys);
        System.out.println("SAMPLING TIME " + (System.nanoTime() - starttime) / 1000000000.0);
    }
}
