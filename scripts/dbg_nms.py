import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import oracle, b200det
from b200det import utils as butils
DEV=torch.device('cuda:0')
T=lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
rng = np.random.default_rng(12)
n, C = 3000, 20
boxes = np.sort(rng.uniform(0, 800, (n, 2, 2)), axis=1).reshape(n, 4)[:, [0, 2, 1, 3]].astype(np.float32)
score = rng.uniform(0, 1, n).astype(np.float32)
label = rng.integers(0, C, n).astype(np.int64)
mx=np.float32(boxes.max()); nb=(boxes+(label.astype(np.float32)*mx)[:,None]).astype(np.float32)
for m in (1000,2000,2048,2049,2500,3000):
    for bx,name in ((boxes,'plain'),(nb,'offset')):
        k=butils.nms(T(bx[:m]),T(score[:m]),0.5).cpu().numpy(); o=oracle.nms(bx[:m],score[:m],0.5)
        print(m,name,len(k),len(o),np.array_equal(k,o), 'unique scores', len(np.unique(score[:m])))
