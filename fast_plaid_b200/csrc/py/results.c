/* CPython helper: [B, K] id / score arrays + per-query counts -> list[list[(int, float)]], the return type of
 * FastPlaid.search (search_on_device's re-zip, fast_plaid.py:247-253).  Building 6 400 tuples for a 64 x 100
 * result in C takes a third of the time of the Python zip, which was the largest host cost of a search call.
 * Host glue only: no search arithmetic lives here. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject* zip_results(PyObject* self, PyObject* args) {
  unsigned long long p_ids, p_scores, p_counts;
  Py_ssize_t B, K;
  if (!PyArg_ParseTuple(args, "KKKnn", &p_ids, &p_scores, &p_counts, &B, &K)) return NULL;
  const int64_t* ids = (const int64_t*)(uintptr_t)p_ids;
  const float* scores = (const float*)(uintptr_t)p_scores;
  const int32_t* counts = (const int32_t*)(uintptr_t)p_counts;
  if (B < 0 || K < 0 || (B > 0 && (!counts || (K > 0 && (!ids || !scores))))) {
    PyErr_SetString(PyExc_ValueError, "zip_results: bad arguments");
    return NULL;
  }
  /* (int, float) tuples cannot form cycles: keep the cyclic GC from scanning the young objects every 700
   * allocations while the lists are built (a third of the time) */
  const int gc_was_enabled = PyGC_Disable();
  PyObject* out = PyList_New(B);
  if (!out) goto fail;
  for (Py_ssize_t b = 0; b < B; ++b) {
    Py_ssize_t n = counts[b];
    if (n < 0) n = 0;
    if (n > K) n = K;
    PyObject* row = PyList_New(n);
    if (!row) goto fail;
    PyList_SET_ITEM(out, b, row);
    for (Py_ssize_t i = 0; i < n; ++i) {
      PyObject* id = PyLong_FromLongLong((long long)ids[b * K + i]);
      PyObject* sc = PyFloat_FromDouble((double)scores[b * K + i]);
      PyObject* t = (id && sc) ? PyTuple_Pack(2, id, sc) : NULL;
      Py_XDECREF(id);
      Py_XDECREF(sc);
      if (!t) goto fail;
      PyList_SET_ITEM(row, i, t);
    }
  }
  if (gc_was_enabled) PyGC_Enable();
  return out;
fail:
  Py_XDECREF(out);
  if (gc_was_enabled) PyGC_Enable();
  return NULL;
}

static PyMethodDef methods[] = {
    {"zip_results", zip_results, METH_VARARGS,
     "zip_results(ids_ptr, scores_ptr, counts_ptr, B, K) -> list[list[(int, float)]] from host int64/float32/int32 arrays"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_fpb_results", NULL, -1, methods};
PyMODINIT_FUNC PyInit__fpb_results(void) { return PyModule_Create(&module); }
