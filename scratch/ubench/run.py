import ctypes, os, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmufu.so"))
L.ubench_main()
