import torch
a = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
for _ in range(6):
    c = a @ b.t()
torch.cuda.synchronize()
a = torch.randn(9472, 3584, device="cuda").to(torch.bfloat16)
b = torch.randn(151936, 3584, device="cuda").to(torch.bfloat16)
for _ in range(3):
    c = a @ b.t()
torch.cuda.synchronize()
