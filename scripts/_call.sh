cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_clinkage.py -m gpu -q --maxfail=6 -p no:cacheprovider > gpurun_out/c19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c19_pytest.log
tail -30 gpurun_out/c19_pytest.log
python - <<'PY' > gpurun_out/c19_time.log 2>&1
import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import hammock_b200 as hb
from hammock_b200 import synth
z = np.load("tests/golden/musi_clinkage.npz")
for rep in range(3):
    t = time.time(); rc, G, err = hb.clinkage_cluster_arrays(z["residues"], z["offsets"], z["abundance"], synth.blosum62(), 20, 3, 0); print("musi clinkage wall ms", (time.time()-t)*1e3, rc)
d = synth.generate(10000, 12, 12, seed=3)
for rep in range(2):
    t = time.time(); rc, G, err = hb.clinkage_cluster_arrays(d["residues"], d["offsets"], d["abundance"], synth.blosum62(), 20, 3, 0); print("10k clinkage wall ms", (time.time()-t)*1e3, rc, len(G.result_order))
PY
cat gpurun_out/c19_time.log
