// Online join-tree planner: C++ mirror of taskdb::planner::Planner
// (/root/reference/prover/crates/taskdb/src/planner/mod.rs:91-240, task.rs:17-80).
// Semantics kept: segments arrive one at a time; two peaks of equal height are joined immediately, so joins
// are created (and numbered) as early as possible; finish() joins the remaining peaks from the smallest up
// and appends Finalize.  Keccak/Union peaks follow the same rule in a separate deque.
#include "../../include/b200zkp.h"
#include <algorithm>
#include <deque>
#include <vector>

struct b200_planner {
    std::vector<b200_task> tasks;
    std::vector<uint32_t> peaks;          // decreasing height
    std::deque<uint32_t> keccak_peaks;    // decreasing height
    size_t consumer_position = 0;
    int64_t last_task = -1;

    uint32_t push(uint32_t command, uint32_t height) {
        b200_task t{};
        t.task_number = (uint32_t)tasks.size();
        t.task_height = height;
        t.command = command;
        tasks.push_back(t);
        return t.task_number;
    }
    uint32_t enqueue_join(uint32_t left, uint32_t right) {
        uint32_t h = 1 + std::max(tasks[left].task_height, tasks[right].task_height);
        uint32_t n = push(B200_CMD_JOIN, h);
        tasks[n].n_depends_on = 2; tasks[n].depends_on[0] = left; tasks[n].depends_on[1] = right;
        return n;
    }
    uint32_t enqueue_union(uint32_t left, uint32_t right) {
        uint32_t h = 1 + std::max(tasks[left].task_height, tasks[right].task_height);
        uint32_t n = push(B200_CMD_UNION, h);
        tasks[n].n_keccak_depends_on = 2; tasks[n].keccak_depends_on[0] = left; tasks[n].keccak_depends_on[1] = right;
        return n;
    }
};

extern "C" {

b200_planner* b200_planner_new(void) { return new b200_planner(); }
void b200_planner_free(b200_planner* pl) { delete pl; }

int64_t b200_planner_enqueue_segment(b200_planner* pl) {
    if (pl->last_task >= 0) return -1;
    uint32_t task_number = pl->push(B200_CMD_SEGMENT, 0);
    uint32_t new_peak = task_number;
    while (!pl->peaks.empty()) {
        uint32_t smallest = pl->peaks.back();
        uint32_t nh = pl->tasks[new_peak].task_height, sh = pl->tasks[smallest].task_height;
        if (nh < sh) break;
        pl->peaks.pop_back();                       // equal heights merge (greater cannot happen)
        new_peak = pl->enqueue_join(smallest, new_peak);
    }
    pl->peaks.push_back(new_peak);
    return task_number;
}

int64_t b200_planner_enqueue_keccak(b200_planner* pl) {
    if (pl->last_task >= 0) return -1;
    uint32_t task_number = pl->push(B200_CMD_KECCAK, 0);
    uint32_t new_peak = task_number;
    while (!pl->keccak_peaks.empty()) {
        uint32_t smallest = pl->keccak_peaks.back();
        uint32_t nh = pl->tasks[new_peak].task_height, sh = pl->tasks[smallest].task_height;
        if (nh < sh) break;
        pl->keccak_peaks.pop_back();
        new_peak = pl->enqueue_union(smallest, new_peak);
    }
    pl->keccak_peaks.push_back(new_peak);
    return task_number;
}

int64_t b200_planner_finish(b200_planner* pl) {
    if (pl->peaks.empty()) return -1;
    // finish unions: fold from the front (highest) pairwise, as the reference does
    bool have_keccak = !pl->keccak_peaks.empty();
    while (pl->keccak_peaks.size() >= 2) {
        uint32_t p0 = pl->keccak_peaks.front(); pl->keccak_peaks.pop_front();
        uint32_t p1 = pl->keccak_peaks.front(); pl->keccak_peaks.pop_front();
        pl->keccak_peaks.push_front(pl->enqueue_union(p1, p0));
    }
    if (pl->last_task < 0) {
        while (pl->peaks.size() >= 2) {
            uint32_t p0 = pl->peaks.back(); pl->peaks.pop_back();
            uint32_t p1 = pl->peaks.back(); pl->peaks.pop_back();
            pl->peaks.push_back(pl->enqueue_join(p1, p0));
        }
        uint32_t dep = pl->peaks[0];
        uint32_t h = 1 + pl->tasks[dep].task_height;
        if (have_keccak) h = std::max(h, 1 + pl->tasks[pl->keccak_peaks[0]].task_height);
        uint32_t n = pl->push(B200_CMD_FINALIZE, h);
        pl->tasks[n].n_depends_on = 1; pl->tasks[n].depends_on[0] = dep;
        if (have_keccak) { pl->tasks[n].n_keccak_depends_on = 1; pl->tasks[n].keccak_depends_on[0] = pl->keccak_peaks[0]; }
        pl->last_task = n;
    }
    return pl->last_task;
}

size_t b200_planner_task_count(const b200_planner* pl) { return pl->tasks.size(); }
int b200_planner_get_task(const b200_planner* pl, size_t i, b200_task* out) {
    if (i >= pl->tasks.size()) return -1;
    *out = pl->tasks[i];
    return 0;
}
int b200_planner_next_task(b200_planner* pl, b200_task* out) {
    if (pl->consumer_position >= pl->tasks.size()) return 1;
    *out = pl->tasks[pl->consumer_position++];
    return 0;
}

}  // extern "C"
