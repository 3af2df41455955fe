// One queued operation of the scheduler classes (wire format of a submit_* call).
// Mirrors reference src/utility/task.h:10-15; on the GPU a batch record is (src, target, value|0).
#pragma once

struct task {
  bool add;    // insert (true) or remove (false) ...
  bool read;   // ... unless this is a neighbourhood read
  int src;
  int target;
};
