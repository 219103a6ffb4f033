"""CPU model of the blend kernels' per-record iteration counts for alternative lane<->pixel layouts, on the bench scenes'
footprint statistics (oracle projection + the record extents of raster_common.cuh::make_record).
  old     : lane = column, 8 rows per lane, slot = one 16x2 row pair  -> visits = row pairs inside the y-extent
  new     : block-cyclic ownership lane = (x mod 8) + 8 (y mod 4) with the per-pixel state in shared memory ->
            iterations = max over lanes of (#columns in box) x (#rows in box)
  altB    : block-cyclic 4 x 8
Used to decide (round 2) NOT to rewrite the kernels around a block-cyclic layout: see profiles/r02_blend_layout_model.md."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from opensplat_b200.scene import make_scene
from oracle import oracle as orc
def sim(n,W,H,scale,opac,label):
    sc=make_scene(n,W,H,scale=scale,sh_degree=0,opacity=opac,seed=0)
    o=orc.project_forward(sc['means'],sc['scales'],1.0,sc['quats'],sc['viewmat'],sc['projmat'],sc['fx'],sc['fy'],sc['cx'],sc['cy'],H,W)
    xy=o['xys'].astype(np.float64); con=o['conics'].astype(np.float64); r=o['radii']; op=sc['opacities'].reshape(-1).astype(np.float64)
    a,b,c=con[:,0],con[:,1],con[:,2]
    smax=np.maximum(0,np.log(op)+5.541263545158426+1e-3)
    det=a*c-b*b
    hx=np.sqrt(2*smax/det*c)*1.001+0.01; hy=np.sqrt(2*smax/det*a)*1.001+0.01
    tx_n,ty_n=(W+15)//16,(H+15)//16
    tcx=xy[:,0]/16; tcy=xy[:,1]/16; tr=r/16.0
    x0=np.clip((tcx-tr).astype(int),0,tx_n); x1=np.clip((tcx+tr+1).astype(int),0,tx_n)
    y0=np.clip((tcy-tr).astype(int),0,ty_n); y1=np.clip((tcy+tr+1).astype(int),0,ty_n)
    tot_pairs=0; kept=0; old_vis=0; new_it=0; box_px=0; new_it_b=0
    hist={}
    for dx in range(0,8):
      for dy in range(0,8):
        ok=(x0+dx<x1)&(y0+dy<y1)&(r>0)
        if not ok.any(): continue
        tx=(x0+dx)[ok]; ty=(y0+dy)[ok]
        gx=xy[ok,0]-16*tx; gy=xy[ok,1]-16*ty
        xlo=np.maximum(np.ceil(gx-hx[ok]),0); xhi=np.minimum(np.floor(gx+hx[ok]),15)
        ylo=np.maximum(np.ceil(gy-hy[ok]),0); yhi=np.minimum(np.floor(gy+hy[ok]),15)
        tot_pairs+=ok.sum()
        k=(xlo<=xhi)&(ylo<=yhi)
        kept+=k.sum()
        xlo,xhi,ylo,yhi=[v[k].astype(int) for v in (xlo,xhi,ylo,yhi)]
        old_vis+=((yhi>>1)-(ylo>>1)+1).sum()
        w=xhi-xlo+1; h=yhi-ylo+1
        box_px+=(w*h).sum()
        # new: per lane counts
        cm=((2<<xhi)-1)&~((1<<xlo)-1); rm=((2<<yhi)-1)&~((1<<ylo)-1)
        nbmax=np.zeros(len(w),int)
        for yl in range(4):
            bm=(rm>>yl)&0x1111
            nb=np.array([bin(v).count('1') for v in bm]) if False else ((bm&1)+((bm>>4)&1)+((bm>>8)&1)+((bm>>12)&1))
            nbmax=np.maximum(nbmax,nb)
        both=np.zeros(len(w),bool)
        for xl in range(8):
            both|=(((cm>>xl)&0x101)==0x101)
        # phase 2 iterations: max nb among lanes with both columns -> approx nbmax
        it=nbmax+np.where(both,nbmax,0)
        new_it+=it.sum()
        # alternative layout B: lane=(x mod 4)+4*(y mod 8): na up to 4, nb up to 2
        # iterations = max over lanes na*nb ~ ceil(w/4)*ceil(h/8)
        new_it_b+=(np.ceil(w/4)*np.ceil(h/8)).sum()
    print(label,'pairs',tot_pairs,'kept',kept, 'kept frac %.3f'%(kept/tot_pairs),'old visits/rec %.2f'%(old_vis/kept),'new iters/rec %.2f'%(new_it/kept),'altB %.2f'%(new_it_b/kept),'box px/rec %.1f'%(box_px/kept), 'lane util new %.2f'%(box_px/(32*new_it)), 'old %.2f'%(box_px/(32*old_vis)))
sim(200000,1920,1080,0.02,(0.05,0.95),'C2-like')   # same footprints; density lower (irrelevant for per-record stats)
sim(200000,2560,1440,0.023,(0.05,0.95),'C5-like')
