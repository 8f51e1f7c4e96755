"""SURVEY.md 8(f) N3 / N4 -- MEASURE (not build) the stages either side of the greedy path at the 1 M scale, on the host
CPU of whatever machine runs this script:

  N4  input side: the C++ host (hammock_greedy --time-host --host-only): fasta parsing + de-duplication, label order,
      automatic parameters, UniqueSequence.sortSequences (UniqueSequence.java:176-203, FileIOManager.java:159-255)
  N2  output side of the host: the four result files of the greedy stage at 1 M, written by the C++ writers from the
      oracle's S1M clustering (tests/cpp/writers_harness.cpp: the functions the driver runs after the GPU call)
  N3  output side: the reference's post-greedy MSA step (Hammock.java:414-426, ClustalRunner.java:113-160): one
      `clustalo -i <fa> -o <aln> --force --wrap=999999` process per multi-member cluster, timed on a sample of the
      clusters of the S1M result (the full oracle run kept by scripts/make_s1m_digest.py) and extrapolated to all of them

usage: python scripts/measure_neighbours.py [sample_clusters=500] [threads=nproc]   -> profiles/r02_neighbour_stages.json
"""
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hammock_b200 import build as hb_build, synth          # noqa: E402

CLUSTALO = "/root/reference/clustal-omega-1.2.0/clustalO-64bit"


def main():
    nsample = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    out = {"machine": {"nproc": os.cpu_count(), "note": "build container CPU, not the GPU box"}}
    d = synth.generate(1000000, 12, 12)
    strs = synth.to_strings(d["residues"], d["offsets"])
    with tempfile.TemporaryDirectory() as tmp:
        # ---- N4: a fasta in shuffled order (so the sort has work to do), one label
        rng = np.random.default_rng(1)
        fa = os.path.join(tmp, "s1m.fa")
        t = time.time()
        with open(fa, "w") as f:
            for k, i in enumerate(rng.permutation(len(strs))):
                f.write(f">{k}|{int(d['abundance'][i])}|lib\n{strs[i]}\n")
        exe = hb_build.build_host()
        r = subprocess.run([exe, "greedy", "-i", fa, "--time-host", "--host-only", "--dump-prepared"], capture_output=True, text=True)
        stages = {}
        for line in r.stderr.splitlines():
            if line.startswith("host time "):
                k, v = line[len("host time "):].rsplit(":", 1)
                stages[k.strip()] = float(v.replace("ms", ""))
        out["N4_input_side_cpp_host"] = {"what": "hammock_greedy --time-host --host-only on a shuffled 1 M-sequence fasta (28 MB), one thread",
                                         "ms": stages, "total_ms": round(sum(stages.values()), 1), "rc": r.returncode}
        out["N4_input_side_cpp_host"]["before_round_2_rewrite_ms"] = {"load + de-duplicate": 3054.64, "labels + automatic parameters": 159.714,
                                                                      "sortSequences": 1431.92, "total": 4646.3}
        z = np.load(os.path.join(ROOT, "scratch", "s1m_oracle_result.npz"))
        cid, rank = z["cluster_id"], z["member_rank"]
        # ---- N2: the writers at 1 M
        harness = os.path.join(tmp, "writers_harness")
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-o", harness, os.path.join(ROOT, "tests", "cpp", "writers_harness.cpp")])
        rt = os.path.join(tmp, "result.txt")
        with open(rt, "w") as f:
            f.write(f"{len(cid)} {len(z['result_order'])}\n")
            for a in (cid, rank, z["result_order"]):
                f.write(" ".join(map(str, a.tolist())) + "\n")
        os.mkdir(os.path.join(tmp, "out"))
        r = subprocess.run([harness, fa, "size", "42", rt, os.path.join(tmp, "out") + "/"], capture_output=True, text=True,
                           env=dict(os.environ, HARNESS_TIMES="1"))
        stages = {}
        for line in r.stderr.splitlines():
            if line.startswith("host time "):
                k, v = line[len("host time "):].rsplit(":", 1)
                stages[k.strip()] = float(v.replace("ms", ""))
        wr = {k: v for k, v in stages.items() if k.startswith(("rebuild", "save", "Save"))}
        out["N2_output_side_cpp_host"] = {"what": "rebuildClusters + the four result files (86 MB) from the oracle's S1M clustering, one thread",
                                          "ms": wr, "total_ms": round(sum(wr.values()), 1), "rc": r.returncode,
                                          "before_round_2_rewrite_total_ms": 1857.5}
        # ---- N3
        order = np.argsort(cid, kind="stable")
        cs = cid[order]
        starts = np.nonzero(np.r_[True, cs[1:] != cs[:-1]])[0]
        ends = np.r_[starts[1:], len(order)]
        multi = [(s, e) for s, e in zip(starts, ends) if e - s > 1]
        sizes = np.array([e - s for s, e in multi])
        pick = rng.choice(len(multi), min(nsample, len(multi)), replace=False)
        jobs = []
        for j in pick:
            s, e = multi[j]
            mem = order[s:e]
            mem = mem[np.argsort(rank[mem])]
            p = os.path.join(tmp, f"{int(cs[s])}.fa")
            with open(p, "w") as f:      # Cluster.getFastaString: >id_k (Cluster.java:167-176)
                for k, m in enumerate(mem):
                    f.write(f">{int(cs[s])}_{k + 1}\n{strs[m]}\n")
            jobs.append((p, p[:-3] + ".aln", e - s))

        def run(job):
            t0 = time.time()
            rc = subprocess.run([CLUSTALO, "-i", job[0], "-o", job[1], "--force", "--wrap=999999"], capture_output=True).returncode
            return time.time() - t0, rc, job[2]
        t = time.time()
        with ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(run, jobs))
        wall = time.time() - t
        per = np.array([r_[0] for r_ in res])
        out["N3_post_greedy_msa"] = {
            "what": f"clustalo 1.2.0 (the binary bundled with the reference), one process per multi-member cluster, {len(jobs)} sampled "
                    f"clusters of the S1M result on {threads} threads (the reference uses its -t thread pool the same way)",
            "clusters_in_result": len(multi), "sampled": len(jobs), "failed": int(sum(1 for r_ in res if r_[1] != 0)),
            "members_per_cluster_mean": float(sizes.mean()), "members_per_cluster_max": int(sizes.max()),
            "sample_members_mean": float(np.mean([r_[2] for r_ in res])),
            "process_seconds_mean": float(per.mean()), "process_seconds_p95": float(np.percentile(per, 95)),
            "sample_wall_s": round(wall, 2),
            "extrapolated_wall_s_all_clusters_same_threads": round(wall * len(multi) / len(jobs), 1),
            "extrapolated_cpu_s_all_clusters": round(float(per.mean()) * len(multi), 1)}
    path = os.path.join(ROOT, "profiles", "r02_neighbour_stages.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
