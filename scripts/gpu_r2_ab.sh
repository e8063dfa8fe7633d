mkdir -p gpurun_out
for rep in 1 2; do for lib in ab_libs/shade_mb4.so ab_libs/shade_mb6.so ab_libs/shade_mb8.so ab_libs/shade_mb10.so; do echo -n "$lib: "; FOUNDATION_PT_LIB=$lib timeout 300 python scripts/probe_render.py terrain 32 2>&1 | tail -1; done; done | tee gpurun_out/r2f_ab_shade.log
for lib in ab_libs/shade_mb4.so ab_libs/shade_mb6.so ab_libs/shade_mb8.so; do for scn in spheres instanced; do echo -n "$lib: "; FOUNDATION_PT_LIB=$lib timeout 300 python scripts/probe_render.py $scn 16 2>&1 | tail -1; done; done | tee -a gpurun_out/r2f_ab_shade.log
