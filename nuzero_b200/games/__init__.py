"""Host-side mirrors of the reference's Game classes (Games/Game.py duck type) backed by the device kernels."""
