"""What an EXACT winner-only vote could save at cfg-2: score all hypotheses on a prefix of the pixel\ntiles, then keep only hypotheses with count_so_far + pixels_left >= best_so_far.  CPU simulation on the bench workload\n(numpy cos test, 2 images x 9 keypoints x 512 hypotheses).  Output: profiles/r02_winner_only_sim.txt."""
import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
from clean_pvnet_b200 import synth
import pvnet_oracle as po
mask, vertex, _ = synth.make_inputs("cfg2", device="cpu", seed=1236, B=2)
m, v = mask.numpy(), vertex.numpy()
out, dbg = po.ransac_voting_layer_v3(m, v, 512, inlier_thresh=0.99, seed=1000, debug=True)
res=[]
for b in range(2):
    # reproduce thinning? use oracle's tn but we do not have the selection; approximate with a random third of fg in raster order
    yx = np.argwhere(m[b]!=0); rng=np.random.default_rng(b); keep = rng.random(len(yx)) < 30000/len(yx)
    yx = yx[keep]; tn=len(yx)
    coords = yx[:, ::-1].astype(np.float32)
    direct = v[b][yx[:,0], yx[:,1]]   # [tn,K,2]
    for k in range(9):
        hyp = dbg["hyp"][b,k]        # [hn,2] (from the true selection; fine as random hypotheses of the right distribution)
        d = hyp[:,None,:]-coords[None,:,:]
        n2 = np.sqrt((d*d).sum(-1)); vv=direct[:,k,:]; n1=np.sqrt((vv*vv).sum(-1))
        cos = (d*vv[None]).sum(-1)/(n2*n1[None]+1e-30)
        inl = cos>0.99
        cum = np.cumsum(inl,axis=1)   # [hn,tn]
        tot = cum[:,-1]
        rho = tot/tn
        best = rho.max()
        def work(fr):
            alive = np.ones(512,bool); w=0; prev=0
            for f in fr:
                w += (f-prev)*alive.mean(); prev=f
                t=int(f*tn)-1
                c=cum[:,t]; bst=c[alive].max()
                alive &= (c + (tn-1-t) >= bst)
            w += (1-prev)*alive.mean()
            return w, alive.sum()
        r={}
        for name,fr in [('1@.75',[.75]),('2',[.6,.8]),('3',[.5,.7,.85]),('4',[.5,.65,.8,.9]),('6',[.4,.5,.6,.7,.8,.9]),('10',[i/20 for i in range(8,20)])]:
            r[name]=work(fr)
        res.append((b,k,best,np.median(rho),(rho<best-1/3).mean(),r))
for x in res: print(x[0],x[1],'best %.3f med %.3f P(rho<best-1/3) %.2f'%(x[2],x[3],x[4]), {k:'%.3f/%d'%v for k,v in x[5].items()})
import statistics
for name in res[0][5]:
    print(name, 'mean work %.3f'%np.mean([x[5][name][0] for x in res]))
