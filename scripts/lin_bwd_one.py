import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import functional as SF
n, f, c = 1632803, 65, 32
gh = torch.randn(n, 32, device="cuda"); x = torch.randn(n, f, device="cuda")
for _ in range(3): SF.lin_bwd(gh, x, c)
torch.cuda.synchronize(); print("ok")
