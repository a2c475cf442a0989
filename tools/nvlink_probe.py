import pynvml, torch, traceback
pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
print("devices", n, "driver", pynvml.nvmlSystemGetDriverVersion())
try:
    print("torch uuid", torch.cuda.get_device_properties(0).uuid)
except Exception as e:
    print("torch uuid err", e)
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    print(i, pynvml.nvmlDeviceGetUUID(h))
    for fid, name in ((138, "DATA_TX"), (139, "DATA_RX"), (136, "RAW_TX"), (137, "RAW_RX")):
        for scope in (0xFFFFFFFF, 0):
            try:
                v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, scope)])[0]
                print("  field", name, "scope", hex(scope), "ret", v.nvmlReturn, "type", v.valueType, "ull", v.value.ullVal, "ul", v.value.ulVal)
            except Exception as e:
                print("  field", name, hex(scope), "EXC", repr(e))
    try:
        print("  link0 state", pynvml.nvmlDeviceGetNvLinkState(h, 0))
    except Exception as e:
        print("  link state EXC", repr(e))
x = torch.ones(64 << 20, device="cuda:0")
y = x.to("cuda:1"); torch.cuda.synchronize()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
v = pynvml.nvmlDeviceGetFieldValues(h, [(138, 0xFFFFFFFF), (139, 0xFFFFFFFF)])
print("after 256 MB copy 0->1:", [(t.nvmlReturn, t.value.ullVal) for t in v])
