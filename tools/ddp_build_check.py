"""2-rank DatasetBuilder check: outputs must equal the single-process golden manifest."""
import hashlib, json, os, sys, tempfile, shutil
sys.path.insert(0, '/root/repo')
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
os.environ["MDB_FRAMES_PER_BATCH"] = "1"
from mddatasetbuilder_b200 import DatasetBuilder
GOLD = '/root/repo/tests/golden'
g = json.load(open(os.path.join(GOLD, 'golden.json')))
work = '/tmp/ddp_work'
ok = True
# the clustered run on ONE rank first (no process group yet): what two ranks must reproduce
solo = tempfile.mkdtemp(prefix='ddp_solo_%d_' % rank)
os.chdir(solo)
b0 = DatasetBuilder(atomname=g["atomname"], bondfilename=os.path.join(GOLD, 'traj', 'bonds.small'),
                    dumpfilename=os.path.join(GOLD, 'traj', 'dump.small'), dataset_name="clus", stepinterval=1, n_clusters=10, seed=3)
b0.builddataset(writegjf=False)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank == 0:
    shutil.rmtree(work, ignore_errors=True); os.makedirs(work)
dist.barrier(device_ids=[local])
os.chdir(work)
b = DatasetBuilder(atomname=g["atomname"], bondfilename=os.path.join(GOLD, 'traj', 'bonds.small'),
                   dumpfilename=os.path.join(GOLD, 'traj', 'dump.small'), dataset_name="gold", cutoff=5.0, stepinterval=2,
                   n_clusters=10000, nproc=1, qmkeywords="%nproc=4\n#force mn15/6-31g** Geom=PrintInputOrient")
b.builddataset()
dist.barrier(device_ids=[local])
if rank == 0:
    got = {}
    for base in (b.dataset_dir, b.gjfdir):
        for dp, _, files in os.walk(base):
            for fn in files:
                p = os.path.join(dp, fn); got[os.path.relpath(p, work)] = hashlib.sha256(open(p, 'rb').read()).hexdigest()
    want = g["e2e_bonds"]["manifest"]
    print("types equal", b.atombondtype == g["e2e_bonds"]["atombondtype"], "nstructure", b._nstructure, g["e2e_bonds"]["nstructure"])
    print("files", len(got), len(want), "identical", got == want)
    ok = ok and got == want and b.atombondtype == g["e2e_bonds"]["atombondtype"]
dist.barrier(device_ids=[local])
# dump-only mode (perceived bond orders in the keys), fragments and atom_pref
b3 = DatasetBuilder(atomname=g["atomname"], dumpfilename=os.path.join(GOLD, 'traj', 'dump.small'), dataset_name="golddump",
                    cutoff=4.0, stepinterval=3, n_clusters=10000, nproc=1, fragment=True, atom_pref=True,
                    qmkeywords=["%nproc=4\n#force mn15/6-31g**", "%nproc=4\n#force mn15/def2tzvp geom=chk"])
b3.builddataset()
dist.barrier(device_ids=[local])
if rank == 0:
    import numpy as np
    got = {}
    for base in (b3.dataset_dir, b3.gjfdir):
        for dp, _, files in os.walk(base):
            for fn in files:
                p = os.path.join(dp, fn)
                got[os.path.relpath(p, work)] = np.load(p).tolist() if fn.endswith(".npy") else hashlib.sha256(open(p, 'rb').read()).hexdigest()
    want = g["e2e_dump"]["manifest"]
    print("dump-only: types equal", b3.atombondtype == g["e2e_dump"]["atombondtype"], "nstructure", b3._nstructure, g["e2e_dump"]["nstructure"],
          "files", len(got), len(want), "identical", got == want)
    ok = ok and got == want
dist.barrier(device_ids=[local])
# clustering path with collectives (k-means partial-sum all-reduce, scaler bounds, selection gather)
b2 = DatasetBuilder(atomname=g["atomname"], bondfilename=os.path.join(GOLD, 'traj', 'bonds.small'),
                    dumpfilename=os.path.join(GOLD, 'traj', 'dump.small'), dataset_name="clus", stepinterval=1, n_clusters=10, seed=3)
b2.builddataset(writegjf=False)
dist.barrier(device_ids=[local])
if rank == 0:
    nx = sum(len(f) for _, _, f in os.walk(b2.dataset_dir))
    print("clustered run: nstructure", b2._nstructure, "xyz files", nx, {t: len(v) for t, v in b2._chosen.items()})
    same = b2._chosen == b0._chosen and b2.atombondtype == b0.atombondtype
    print("two ranks select what one rank selects:", same)
    ok = ok and same and nx == b2._nstructure
t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0 and int(t.item()) == 1:
    print("DDP-BUILD-OK")
print("rank", rank, "done")
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
