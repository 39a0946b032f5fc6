#pragma once
#include <signal.h>
#include <ucontext.h>
#include <stdio.h>
#include <unistd.h>
#include <string.h>
static void b2_on_ill(int, siginfo_t *, void *uc) { fprintf(stderr, "SIGILL at %llx\n", (unsigned long long)((ucontext_t *)uc)->uc_mcontext.gregs[REG_RIP]); _exit(99); }
__attribute__((constructor)) static void b2_inst_ill() { struct sigaction sa; memset(&sa, 0, sizeof(sa)); sa.sa_sigaction = b2_on_ill; sa.sa_flags = SA_SIGINFO; sigaction(SIGILL, &sa, nullptr); }
