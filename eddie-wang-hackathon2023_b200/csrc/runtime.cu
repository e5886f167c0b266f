// runtime.cu -- error plumbing, launch counter, device checks for the C ABI.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace b200
{

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n)
{
    g_launches.fetch_add(static_cast<unsigned long long>(n), std::memory_order_relaxed);
}

static int g_sms = 0;
static int g_dev_ok = -1;
static std::mutex g_mu;

static void probe()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_dev_ok >= 0)
        return;
    int dev = 0, major = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess
        || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess
        || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    {
        cudaGetLastError();
        g_dev_ok = 0;
        return;
    }
    g_sms = sms;
    g_dev_ok = (major == 10) ? 1 : 0;
}

bool device_ok()
{
    if (g_dev_ok < 0)
        probe();
    return g_dev_ok == 1;
}

static std::atomic<int> g_pdl{-1};

bool pdl_enabled()
{
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0)
    {
        const char* e = getenv("B200_PDL");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v == 1;
}

// Per calling thread: the hint describes the caller's OWN launch sequence (its KV caches are not written by the
// kernel in front of the attention launch), so it must not leak into launches other threads / plugin contexts make.
static thread_local int g_static_kv = 0;

bool static_kv_hint()
{
    return g_static_kv == 1;
}

void set_static_kv(int on)
{
    g_static_kv = on ? 1 : 0;
}

void set_pdl(int on)
{
    g_pdl = on ? 1 : 0;
}

int num_sms()
{
    if (g_dev_ok < 0)
        probe();
    return g_sms > 0 ? g_sms : 148;
}

} // namespace b200

extern "C"
{

const char* b200_last_error(void)
{
    return b200::g_err;
}

int b200_abi_version(void)
{
    return 1;
}

unsigned long long b200_launch_count(void)
{
    return b200::g_launches.load(std::memory_order_relaxed);
}

int b200_set_pdl(int enabled)
{
    b200::set_pdl(enabled);
    return B200_OK;
}

int b200_set_static_kv_hint(int enabled)
{
    b200::set_static_kv(enabled);
    return B200_OK;
}
}
