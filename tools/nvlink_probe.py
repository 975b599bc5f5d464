"""What NVLink traffic counters does this box expose?  (run under gpurun; diagnostic for bench.py's `nvlink` field)"""
import subprocess
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
names = [n for n in dir(pynvml) if n.startswith("NVML_FI_DEV_NVLINK_THROUGHPUT")]
print(names)
for n in names:
    fid = getattr(pynvml, n)
    for scope in (None, 0, 0xFFFFFFFF):
        try:
            v = pynvml.nvmlDeviceGetFieldValues(h, [fid if scope is None else (fid, scope)])[0]
            print(n, "scope", scope, "ret", v.nvmlReturn, "type", v.valueType, "ull", v.value.ullVal)
        except Exception as e:
            print(n, "scope", scope, "EXC", e)
print(subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", "0"], capture_output=True, text=True).stdout[:1500])
