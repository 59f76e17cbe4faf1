import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import rtb200 as rt
for name,(w,h,spp) in {"cornell_box":(600,600,32),"random_scene":(400,225,100),"final_scene":(800,800,16),"obj_cornell":(600,600,16)}.items():
    s=rt.Scene(seed=2022).build_config(name, w/h)
    gpu=rt.GpuScene(s.lower(),0)
    gpu.render(w,h,0,2)
    t=time.time()
    acc,st=gpu.render(w,h,0,spp)
    dt=time.time()-t
    print(name, f"rays {st.rays} samples {st.samples} ms_total {st.ms_total:.1f} wall {dt*1e3:.1f} extend {st.ms_extend:.1f} shade {st.ms_shade:.1f} other {st.ms_other:.1f} launches {st.launches} Mrays/s {st.rays/st.ms_total/1e3:.1f} rays/sample {st.rays/st.samples:.2f}")
    print("  per bounce", list(st.rays_per_bounce[:12]), "...", st.rays_per_bounce[49])
    img=gpu.resolve(acc,w,h,spp)
    from PIL import Image
    Image.fromarray(img).save(f"gpurun_out/{name}.png")
