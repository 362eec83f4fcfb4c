/* Diagnosis aid (not part of the product): a SIGABRT handler that writes the C/C++ stack of the aborting
 * thread to stderr.  Used to find which library's static destruction calls std::terminate at interpreter exit
 * (bench.py, D4_BENCH_ABORT_TRACE=<path of the built .so>).  Build: gcc -shared -fPIC -O1 -o abort_trace.so abort_trace.c */
#include <execinfo.h>
#include <signal.h>
#include <string.h>
#include <unistd.h>

static void on_abort(int sig) {
    static const char head[] = "\n== abort_trace: SIGABRT, stack of the aborting thread ==\n";
    void* frames[96];
    (void)sig;
    (void)!write(2, head, sizeof(head) - 1);
    backtrace_symbols_fd(frames, backtrace(frames, 96), 2);
    _exit(134);
}

void abort_trace_install(void) {
    void* warm[4];
    backtrace(warm, 4); /* loads libgcc now, not inside the handler */
    signal(SIGABRT, on_abort);
}
