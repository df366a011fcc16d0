"""CPU oracle of the x4 SR hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package; the product under image-enhance-keras_b200/ never does and has no CPU path.

Pinned against the reference's own code (oracle/refgen.py -> tests/golden/*.npz): tiling.py, psnr part of
scoring.py, imgpatch restatement, cv2 colour.  PARITY UNPINNED (no reference implementation or golden
vector available offline): model.py (Keras/TF conv stack, bilinear, Adam), scoring.py SSIM / rgb2ycbcr
(skimage), shuffle.py (tf.depth_to_space / Theano branches).
"""
