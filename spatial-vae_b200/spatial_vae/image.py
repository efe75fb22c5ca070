"""Image-stack helpers used by the particle CLI (host preprocessing; reference spatial_vae/image.py)."""


def crop(stack, size):
    """Centre crop of the last two axes to size x size (reference image.py:32-44)."""
    n, m = stack.shape[-2:]
    top, left = (n - size) // 2, (m - size) // 2
    return stack[..., top:top + size, left:left + size]
