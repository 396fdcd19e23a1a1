mkdir -p gpurun_out
( lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; nvidia-smi topo -m 2>&1 | head -20; free -g | head -2; python - <<'P'
import os
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    print("nvml cpu affinity", list(pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)))
except Exception as e:
    print("pynvml:", e)
P
) 2>&1 | tee gpurun_out/topo_ay.log
