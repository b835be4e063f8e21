/* Host-side marshalling helper for ReplayDataset.extend (agent0/deepq/replay.py:45-53): walks the reference's
 * list of (blob, action, reward, done) tuples with the CPython API and fills plain C arrays, so that the 1280
 * entries of a Trainer.step cost ~40 us instead of ~400 us of zip / np.asarray / b"".join in Python.  Nothing
 * here computes anything: the pointers it collects are handed to the C ABI (a0_ex_extend_v), which copies the
 * compressed bytes straight into its page-locked staging block.  Loaded with ctypes.PyDLL (the GIL stays held);
 * built by agent0_b200/build.py with gcc against the interpreter's headers.  Not part of include/agent0_b200.h:
 * it is Python-specific glue, the C ABI below it is what a binding in another language would call.            */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

/* transitions: list or tuple of 4-tuples.  blob: bytes, bytearray or any object with the buffer protocol
 * (C-contiguous).  Returns the number of entries unpacked, or -1 with a Python exception set.
 * views: caller-allocated Py_buffer[m] (released by a0_py_release).                                          */
int64_t a0_py_unpack(PyObject* transitions, int64_t m, const uint8_t** ptrs, int64_t* lens, int64_t* action,
                     double* reward, uint8_t* done, Py_buffer* views) {
  PyObject* seq = PySequence_Fast(transitions, "extend() expects a list of (frames, action, reward, done) tuples");
  if (!seq) return -1;
  if (PySequence_Fast_GET_SIZE(seq) != m) {
    Py_DECREF(seq);
    PyErr_SetString(PyExc_ValueError, "extend(): the list changed size");
    return -1;
  }
  int64_t i = 0;
  for (; i < m; ++i) {
    PyObject* t = PySequence_Fast_GET_ITEM(seq, i);
    if (!PyTuple_Check(t) || PyTuple_GET_SIZE(t) != 4) {
      PyErr_Format(PyExc_TypeError, "extend(): entry %lld is not a (frames, action, reward, done) tuple", (long long)i);
      break;
    }
    PyObject* blob = PyTuple_GET_ITEM(t, 0);
    views[i].obj = NULL;
    if (PyBytes_CheckExact(blob)) {
      ptrs[i] = (const uint8_t*)PyBytes_AS_STRING(blob);
      lens[i] = (int64_t)PyBytes_GET_SIZE(blob);
    } else {
      if (PyObject_GetBuffer(blob, &views[i], PyBUF_C_CONTIGUOUS) != 0) break;
      ptrs[i] = (const uint8_t*)views[i].buf;
      lens[i] = (int64_t)views[i].len;
    }
    PyObject* a = PyTuple_GET_ITEM(t, 1);
    long long av = PyLong_Check(a) ? PyLong_AsLongLong(a) : -1;
    if (!PyLong_Check(a)) {
      PyObject* ai = PyNumber_Index(a);
      if (!ai) { i++; break; }
      av = PyLong_AsLongLong(ai);
      Py_DECREF(ai);
    }
    if (av == -1 && PyErr_Occurred()) { i++; break; }
    action[i] = (int64_t)av;
    const double rv = PyFloat_AsDouble(PyTuple_GET_ITEM(t, 2));
    if (rv == -1.0 && PyErr_Occurred()) { i++; break; }
    reward[i] = rv;
    const int dv = PyObject_IsTrue(PyTuple_GET_ITEM(t, 3));
    if (dv < 0) { i++; break; }
    done[i] = (uint8_t)dv;
  }
  Py_DECREF(seq);
  if (i < m) {
    for (int64_t k = 0; k < i; ++k)
      if (views[k].obj) PyBuffer_Release(&views[k]);
    return -1;
  }
  return m;
}

void a0_py_release(int64_t m, Py_buffer* views) {
  for (int64_t k = 0; k < m; ++k)
    if (views[k].obj) PyBuffer_Release(&views[k]);
}

int64_t a0_py_buffer_size(void) { return (int64_t)sizeof(Py_buffer); }
