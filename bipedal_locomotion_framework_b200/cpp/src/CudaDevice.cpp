/**
 * @file CudaDevice.cpp
 * RAII owner of a C-ABI handle (include/blf_ccm.h).  The only translation units that see the C
 * ABI are this one, ContinuousContactModel.cpp, ContinuousContactModelBatch.cpp and DeviceSoA.cpp.
 */
#include <cstdlib>
#include <iostream>

#include <BipedalLocomotion/ContactModels/ContinuousContactModelBatch.h>

#include "blf_ccm.h"

using namespace BipedalLocomotion::ContactModels;

struct CudaDevice::Impl
{
    blf_ccm_handle* handle{nullptr};
    int index{-1};
};

CudaDevice::CudaDevice() : m_impl(std::make_unique<Impl>()) {}

CudaDevice::~CudaDevice()
{
    if (m_impl && m_impl->handle) blf_ccm_destroy(m_impl->handle);
}

std::shared_ptr<CudaDevice> CudaDevice::open(int device)
{
    std::shared_ptr<CudaDevice> dev(new CudaDevice());
    const int rc = blf_ccm_create(device, &dev->m_impl->handle);
    if (rc != BLF_CCM_OK)
    {
        std::cerr << "[CudaDevice::open] Unable to open CUDA device " << device << ": "
                  << blf_ccm_last_error() << std::endl;
        return nullptr;
    }
    dev->m_impl->index = device;
    return dev;
}

namespace
{
int g_defaultIndex = -1;
}

int CudaDevice::defaultIndex()
{
    if (g_defaultIndex >= 0) return g_defaultIndex;
    if (const char* env = std::getenv("BLF_CCM_DEVICE")) return std::atoi(env);
    return 0;
}

void CudaDevice::setDefaultIndex(int device) { g_defaultIndex = device; }

void* CudaDevice::handle() const { return m_impl->handle; }
int CudaDevice::index() const { return m_impl->index; }
std::int64_t CudaDevice::kernelLaunches() const { return blf_ccm_launch_count(m_impl->handle); }
const char* CudaDevice::lastError() const { return blf_ccm_last_error(); }
