"""Which stage of FOURIER_INV is not run-to-run reproducible at 24 complex slices: torch.fft (cuFFT) calls of the shapes the
method issues, repeated on the same input with the allocator state changed in between."""
import torch

torch.manual_seed(0)
for shape, fn, name in (((24, 50, 256), lambda x: torch.fft.fft(x, dim=2), "fft 256, batch 24 x 50"),
                        ((24, 50, 256), lambda x: torch.fft.ifft(x, dim=2, norm="forward"), "ifft 256, batch 24 x 50"),
                        ((24, 50, 64), lambda x: torch.fft.fft(x, dim=-1), "fft 64, batch 24 x 50"),
                        ((24, 128, 128), lambda x: torch.fft.ifft2(x, dim=(-2, -1), norm="forward"), "ifft2 128^2, batch 24"),
                        ((16, 128, 128), lambda x: torch.fft.ifft2(x, dim=(-2, -1), norm="forward"), "ifft2 128^2, batch 16"),
                        ((40, 160, 160), lambda x: torch.fft.ifft2(x, dim=(-2, -1), norm="forward"), "ifft2 160^2, batch 40")):
    x = torch.randn(shape, dtype=torch.complex64, device="cuda")
    ref = fn(x)
    same = True
    for k in range(6):
        junk = torch.empty(1000 * (k + 1) + 13, device="cuda")  # shift the allocator's addresses
        y = fn(x.clone())
        same = same and torch.equal(ref, y)
        del junk
    print(f"{name}: reproducible = {same}")
