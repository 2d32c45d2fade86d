/* TEST INFRASTRUCTURE ONLY. Definitions of the handful of process-level symbols the reference's
 * PHY sources expect from the -rdynamic softmodem executable (reference CMakeLists.txt:164,
 * SURVEY.md section 8b "loader gotchas" ii).  Linked into every oracle/_ref/libref_*.so with
 * hidden-from-nobody default visibility; they do nothing. */
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
int opp_enabled = 0;
static char g_log_storage[1 << 20];
void *g_log = g_log_storage;
void logRecord_mt(const char *file, const char *func, int line, int comp, int level, const char *fmt, ...) { (void)file; (void)func; (void)line; (void)comp; (void)level; (void)fmt; }
void exit_function(const char *file, const char *function, const int line, const char *s, const int assert_)
{ fprintf(stderr, "reference exit_function: %s:%d %s %s\n", file, line, function, s ? s : ""); (void)assert_; abort(); }
int write_file_matlab(const char *fname, const char *vname, void *data, int length, int dec, unsigned int format, int multiVec) { (void)fname; (void)vname; (void)data; (void)length; (void)dec; (void)format; (void)multiVec; return 0; }
