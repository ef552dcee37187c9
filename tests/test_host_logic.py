"""Host-side logic that needs no GPU (and no compute call into the C-ABI library)."""


def test_plan_lanes_and_sync_marks():
    """Two-lane plan bookkeeping (engine.Plan): lane of every step and the fork / join edges, no GPU needed."""
    from rdfc_gan_b200.engine import Plan
    p = Plan()

    def add(name):
        p.steps.append(lambda s: None)
        p.names.append(name)

    add('stems')
    p.fork()
    p.lane(0); add('r.en2'); add('r.en3')
    p.lane(1); add('d.en2'); add('d.en3')
    p.join()
    p.fork()
    add('fuse1'); add('r.de5')
    p.lane(1); add('d.de5')
    p.join()
    add('fuse_depth')
    assert p.lanes() == [0, 0, 0, 1, 1, 0, 0, 1, 0]
    assert p.sync == {1: ['fork'], 5: ['join', 'fork'], 8: ['join']}
