"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU (numpy / scipy, fp64) restatement of the reference's patch-wise 3D correspondence
and rigid-estimation path.  Every function cites the reference file:line it follows
(paths relative to the gseg-ethz/fusion4landslide tree).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package, and only as the checker / the CPU arm.  The product
package `fusion4landslide_b200` never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * rigid.py (Kabsch / Procrustes / residual filters)  -- PINNED against the reference's own
    functions imported unmodified through oracle/ref_shim.py (tests/golden/rigid_*.npz).
  * knn.py, desc_nn.py                                  -- PINNED against the reference's own
    calls (sklearn kd_tree, scipy cKDTree, torch.cdist+min) on the golden inputs.
  * icp.py (Open3D 0.19 registration_icp), octree part of piecewise.py (Open3D Octree),
    hnswlib / faiss HNSW                                -- PARITY UNPINNED: third-party code that
    is neither vendored in the reference nor installed here; restated from the published
    algorithm, anchored on the reference's call sites.  The reference has no tests.
"""
