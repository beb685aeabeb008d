"""Join-tree planner: the reference's own unit tests restated against the C++ mirror
(/root/reference/prover/crates/taskdb/src/planner/mod.rs:254-453)."""
import pytest

from boundless_b200.planner import CMD_FINALIZE, CMD_JOIN, CMD_KECCAK, CMD_SEGMENT, CMD_UNION, Planner, PlannerErr


def test_simple_plan(b200lib):
    planner = Planner()
    assert planner.enqueue_segment() == 0
    task = planner.next_task()
    assert task.keccak_depends_on == [] and task.depends_on == []
    assert task.command == CMD_SEGMENT and task.task_number == 0
    assert planner.next_task() is None
    assert planner.enqueue_keccak() == 1
    task = planner.next_task()
    assert task.command == CMD_KECCAK and task.task_number == 1 and task.depends_on == []
    assert planner.next_task() is None
    planner.finish()
    task = planner.next_task()
    assert task.command == CMD_FINALIZE
    assert len(task.keccak_depends_on) == 1 and task.task_number == 2 and len(task.depends_on) == 1
    assert task.task_height == 1
    assert planner.task_count() == 3
    assert planner.get_task(0).task_number == 0


def test_balanced(b200lib):
    planner = Planner()
    planner.enqueue_segment()
    t = planner.next_task()
    assert (t.command, t.task_number, t.task_height) == (CMD_SEGMENT, 0, 0)
    assert planner.next_task() is None
    planner.enqueue_keccak()
    t = planner.next_task()
    assert (t.command, t.task_number, t.task_height) == (CMD_KECCAK, 1, 0)
    planner.enqueue_segment()
    t = planner.next_task()
    assert (t.command, t.task_number, t.task_height, t.depends_on) == (CMD_SEGMENT, 2, 0, [])
    j = planner.next_task()
    assert (j.command, j.task_number, j.task_height, len(j.depends_on)) == (CMD_JOIN, 3, 1, 2)
    planner.enqueue_keccak()
    t = planner.next_task()
    assert (t.command, t.task_number, t.task_height) == (CMD_KECCAK, 4, 0)
    u = planner.next_task()
    assert (u.command, u.task_number, u.task_height, len(u.keccak_depends_on)) == (CMD_UNION, 5, 1, 2)
    planner.finish()
    last = planner.next_task()
    assert (last.command, last.task_number, last.task_height) == (CMD_FINALIZE, 6, 2)
    assert len(last.depends_on) == 1 and len(last.keccak_depends_on) == 1


def test_unbalanced_keccak(b200lib):
    planner = Planner()
    planner.enqueue_keccak(); planner.enqueue_keccak(); planner.enqueue_keccak()
    planner.enqueue_segment()
    planner.finish()
    expect = [(0, CMD_KECCAK, 0), (1, CMD_KECCAK, 0), (2, CMD_UNION, 1), (3, CMD_KECCAK, 0), (4, CMD_SEGMENT, 0),
              (5, CMD_UNION, 2), (6, CMD_FINALIZE, 3)]
    for num, cmd, h in expect:
        t = planner.next_task()
        assert (t.task_number, t.command, t.task_height) == (num, cmd, h)


def test_unbalanced(b200lib):
    planner = Planner()
    for _ in range(3):
        planner.enqueue_segment()
    planner.finish()
    expect = [(0, CMD_SEGMENT), (1, CMD_SEGMENT), (2, CMD_JOIN), (3, CMD_SEGMENT), (4, CMD_JOIN), (5, CMD_FINALIZE)]
    for num, cmd in expect:
        t = planner.next_task()
        assert (t.task_number, t.command) == (num, cmd)
    assert t.task_height == 3


def test_err_not_started(b200lib):
    with pytest.raises(PlannerErr, match="PlanNotStartedString"):
        Planner().finish()


def test_err_finalized(b200lib):
    planner = Planner()
    planner.enqueue_segment()
    planner.finish()
    with pytest.raises(PlannerErr, match="PlanFinalized"):
        planner.enqueue_segment()


def test_err_bad_task_numb(b200lib):
    with pytest.raises(IndexError, match="Invalid task number 100"):
        Planner().get_task(100)


@pytest.mark.parametrize("n", [1, 2, 5, 8, 31, 256])
def test_tree_shape_invariants(b200lib, n):
    """n segments -> n-1 joins + finalize; joins are created as soon as two equal-height peaks exist (so their task
    numbers interleave with the segments), and the left input always covers earlier segments than the right."""
    planner = Planner()
    for _ in range(n):
        planner.enqueue_segment()
    fin = planner.finish()
    tasks = [planner.get_task(i) for i in range(planner.task_count())]
    assert sum(t.command == CMD_SEGMENT for t in tasks) == n
    assert sum(t.command == CMD_JOIN for t in tasks) == n - 1
    assert tasks[fin].command == CMD_FINALIZE and fin == len(tasks) - 1
    span = {}
    k = 0
    for t in tasks:
        if t.command == CMD_SEGMENT:
            span[t.task_number] = (k, k); k += 1
        elif t.command == CMD_JOIN:
            l, r = t.depends_on
            assert l < t.task_number and r < t.task_number
            assert span[l][1] + 1 == span[r][0]
            span[t.task_number] = (span[l][0], span[r][1])
    assert span[tasks[fin].depends_on[0]] == (0, n - 1)
    if n & (n - 1) == 0 and n > 1:
        assert tasks[fin].task_height == n.bit_length()       # perfect tree: log2(n) joins + finalize
