mkdir -p gpurun_out
for w in 8 16 32; do echo -n "wave_samples=$w: "; FOUNDATION_PT_WAVE_SAMPLES=$w timeout 300 python scripts/probe_render.py terrain 64 2>&1 | tail -1; done | tee gpurun_out/r2i_wave_samples.log
for w in 8 16 32; do echo -n "wave_samples=$w: "; FOUNDATION_PT_WAVE_SAMPLES=$w timeout 300 python scripts/probe_render.py spheres 64 2>&1 | tail -1; done | tee -a gpurun_out/r2i_wave_samples.log
