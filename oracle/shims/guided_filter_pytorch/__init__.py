"""Restatement (test infrastructure) of the un-vendored third-party package
``guided_filter_pytorch`` (PyPI distribution of wuhuikai/DeepGuidedFilter) that the
reference imports at core/model_fusion_auto.py:2 and calls at :529-530.  No version
is pinned by the reference (no requirements file; README.md:32-33 pins only
Python/PyTorch), and the reference has no tests for it: PARITY UNPINNED for this
dependency.  The published algorithm is restated in box_filter.py / guided_filter.py."""
