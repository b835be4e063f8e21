"""CPU oracle for the agent0 deepq replay-and-target path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``agent0_b200/`` may import this package.  It is used by ``tests/``,
by ``__graft_entry__.smoke()`` and by ``bench.py``'s cpu_baseline / ``--impl reference``
legs, and only as the checker or the timed CPU baseline - never on the product path.

Contents
  reference_replay.py  numpy restatement of Actor.sample's n-step packer
                       (agent0/deepq/agent.py:57-81), ReplayDataset
                       (agent0/deepq/replay.py:14-59) and the IS-weight block of
                       Trainer.step (agent0/deepq/trainer.py:88-96)
  losses.py            numpy fp32 restatement of the six train_step rules
                       (agent0/deepq/agent.py:110-119,172-388) with closed-form
                       gradients of (loss*w).sum() w.r.t. the online-network output
  sumtree.py           the fp32 sum-tree the CUDA sampler must match bit-for-bit
                       (no reference counterpart: the reference's intended law is
                       torch.multinomial over the flat priority vector, replay.py:39-43)
  cpu_path.py          the timed CPU baseline: the reference's host pipeline
                       (deque of lz4 blobs -> __getitem__ -> collate -> .float() -> IS
                       weights -> torch-CPU loss -> update_priority), restated

Parity pinning: the reference has no tests, golden vectors or fixtures,
so the pins are outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference unmodified) and committed
under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function here
against them.
"""
