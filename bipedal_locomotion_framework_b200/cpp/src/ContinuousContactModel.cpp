/**
 * @file ContinuousContactModel.cpp
 * Per-instance facade of the continuous contact model.  Every compute*() is ONE evaluation of the
 * batched CUDA path with n = 1 (blf_ccm_eval_batch_host); there is no host arithmetic here.
 * Interface and parameter handling follow the reference
 * (src/ContactModels/src/ContinuousContactModel.cpp:16-77, :256-274).
 */
#include <cstdlib>
#include <cstring>
#include <iostream>

#include <BipedalLocomotion/ContactModels/ContinuousContactModel.h>
#include <BipedalLocomotion/ContactModels/ContinuousContactModelBatch.h>

#include "blf_ccm.h"

using namespace BipedalLocomotion::ContactModels;
using namespace BipedalLocomotion::ParametersHandler;

namespace
{
blf_ccm_handle* raw(const std::shared_ptr<CudaDevice>& d)
{
    return static_cast<blf_ccm_handle*>(d->handle());
}
} // namespace

void ContinuousContactModel::setDefaultDevice(int device) { CudaDevice::setDefaultIndex(device); }

ContinuousContactModel::ContinuousContactModel()
{
    m_controlMatrix.zero();
    m_autonomousDynamics.zero();
    m_regressor.resize(6, 2);
    m_regressor.zero();
}

bool ContinuousContactModel::initializePrivate(std::weak_ptr<IParametersHandler> weakHandler)
{
    m_allValid = false;
    auto handler = weakHandler.lock();
    if (handler == nullptr)
    {
        std::cerr << "[ContinuousContactModel::initialize] The parameter handler is corrupted. "
                     "Please make sure that the handler exists."
                  << std::endl;
        return false;
    }
    if (!handler->getParameter("length", m_length))
    {
        std::cerr << "[ContinuousContactModel::initialize] Unable to get the variable named length."
                  << std::endl;
        return false;
    }
    if (!handler->getParameter("width", m_width))
    {
        std::cerr << "[ContinuousContactModel::initialize] Unable to get the variable named width."
                  << std::endl;
        return false;
    }
    if (!handler->getParameter("spring_coeff", m_springCoeff))
    {
        std::cerr << "[ContinuousContactModel::initialize] Unable to get the variable named "
                     "spring_coeff."
                  << std::endl;
        return false;
    }
    if (!handler->getParameter("damper_coeff", m_damperCoeff))
    {
        std::cerr << "[ContinuousContactModel::initialize] Unable to get the variable named "
                     "damper_coeff."
                  << std::endl;
        return false;
    }
    if (m_device == nullptr)
    {
        m_device = CudaDevice::open(CudaDevice::defaultIndex());
        if (m_device == nullptr)
        {
            std::cerr << "[ContinuousContactModel::initialize] The CUDA backend is not available and "
                         "there is no CPU evaluation path."
                      << std::endl;
            return false;
        }
    }
    return true;
}

void ContinuousContactModel::setNullForceTransformPrivate(const iDynTree::Transform& transform)
{
    m_nullForceTransform = transform;
    m_allValid = false;
}

void ContinuousContactModel::setStatePrivate(const iDynTree::Twist& twist,
                                             const iDynTree::Transform& transform)
{
    m_twist = twist;
    m_frameTransform = transform;
    m_allValid = false;
}

bool ContinuousContactModel::pushParameters()
{
    if (m_device == nullptr)
    {
        std::cerr << "[ContinuousContactModel] initialize() has not succeeded: no CUDA backend, and "
                     "there is no CPU evaluation path. The result is left untouched."
                  << std::endl;
        return false;
    }
    // pushed before every evaluation so that springCoeff()/damperCoeff() written through the
    // mutable reference take effect at the next (re)computation, exactly as in the reference
    return blf_ccm_set_uniform_params(raw(m_device), m_length, m_width, m_springCoeff, m_damperCoeff)
           == BLF_CCM_OK;
}

const double* ContinuousContactModel::evaluateAll()
{
    const double live[4] = {m_length, m_width, m_springCoeff, m_damperCoeff};
    if (m_allValid && std::memcmp(live, m_allParameters, sizeof(live)) == 0) return m_all;
    m_allValid = false;
    if (!pushParameters())
    {
        m_computeFailed = true;
        return nullptr;
    }
    const int rc = blf_ccm_eval_batch_host(raw(m_device), 1,
                                           reinterpret_cast<const double*>(&m_twist),
                                           reinterpret_cast<const double*>(&m_frameTransform),
                                           reinterpret_cast<const double*>(&m_nullForceTransform),
                                           nullptr,
                                           BLF_CCM_WRENCH | BLF_CCM_AUTODYN | BLF_CCM_CTRL | BLF_CCM_REGRESSOR,
                                           m_all, m_all + 6, m_all + 24, m_all + 12);
    if (rc != BLF_CCM_OK)
    {
        std::cerr << "[ContinuousContactModel] evaluation failed: " << blf_ccm_last_error()
                  << std::endl;
        m_computeFailed = true;
        return nullptr;
    }
    std::memcpy(m_allParameters, live, sizeof(live));
    m_allValid = true;
    return m_all;
}

void ContinuousContactModel::computeContactWrench()
{
    if (const double* all = evaluateAll()) std::memcpy(m_contactWrench.data(), all, 6 * sizeof(double));
}

void ContinuousContactModel::computeAutonomousDynamics()
{
    if (const double* all = evaluateAll())
        std::memcpy(m_autonomousDynamics.data(), all + 6, 6 * sizeof(double));
}

void ContinuousContactModel::computeControlMatrix()
{
    if (const double* all = evaluateAll())
        std::memcpy(m_controlMatrix.data(), all + 24, 36 * sizeof(double));
}

void ContinuousContactModel::computeRegressor()
{
    if (const double* all = evaluateAll()) std::memcpy(m_regressor.data(), all + 12, 12 * sizeof(double));
}

namespace
{
bool surfacePoint(const std::shared_ptr<CudaDevice>& dev, const iDynTree::Twist& twist,
                  const iDynTree::Transform& frame, const iDynTree::Transform& nullForce, double x,
                  double y, double* force, double* torque)
{
    blf_ccm_handle* h = raw(dev);
    void* d = nullptr;
    if (blf_ccm_device_alloc(h, 8 * sizeof(double), &d) != BLF_CCM_OK) return false;
    double* dd = static_cast<double*>(d);
    const double xy[2] = {x, y};
    double out[6];
    bool ok = blf_ccm_copy_h2d(h, dd, xy, sizeof(xy), nullptr) == BLF_CCM_OK
              && blf_ccm_eval_surface_points(h, reinterpret_cast<const double*>(&twist),
                                             reinterpret_cast<const double*>(&frame),
                                             reinterpret_cast<const double*>(&nullForce), 1, dd,
                                             dd + 2, dd + 5, nullptr) == BLF_CCM_OK
              && blf_ccm_copy_d2h(h, out, dd + 2, sizeof(out), nullptr) == BLF_CCM_OK
              && blf_ccm_stream_synchronize(h, nullptr) == BLF_CCM_OK;
    blf_ccm_device_free(h, d);
    if (!ok) return false;
    for (int i = 0; i < 3; ++i)
    {
        if (force) force[i] = out[i];
        if (torque) torque[i] = out[3 + i];
    }
    return true;
}
} // namespace

iDynTree::Force ContinuousContactModel::getForceAtPoint(const double& x, const double& y)
{
    iDynTree::Force force;
    force.zero();
    if (!pushParameters()) return force;
    if (!surfacePoint(m_device, m_twist, m_frameTransform, m_nullForceTransform, x, y, force.data(),
                      nullptr))
        std::cerr << "[ContinuousContactModel::getForceAtPoint] " << blf_ccm_last_error() << std::endl;
    return force;
}

iDynTree::Torque ContinuousContactModel::getTorqueGeneratedAtPoint(const double& x, const double& y)
{
    iDynTree::Torque torque;
    torque.zero();
    if (!pushParameters()) return torque;
    if (!surfacePoint(m_device, m_twist, m_frameTransform, m_nullForceTransform, x, y, nullptr,
                      torque.data()))
        std::cerr << "[ContinuousContactModel::getTorqueGeneratedAtPoint] " << blf_ccm_last_error()
                  << std::endl;
    return torque;
}

const double& ContinuousContactModel::springCoeff() const { return m_springCoeff; }
double& ContinuousContactModel::springCoeff() { return m_springCoeff; }
const double& ContinuousContactModel::damperCoeff() const { return m_damperCoeff; }
double& ContinuousContactModel::damperCoeff() { return m_damperCoeff; }

void ContinuousContactModel::batchInputs(double state[30], double parameters[4]) const
{
    static_assert(sizeof(iDynTree::Twist) == 48 && sizeof(iDynTree::Transform) == 96, "C-ABI row layout");
    std::memcpy(state, &m_twist, 48);
    std::memcpy(state + 6, &m_frameTransform, 96);
    std::memcpy(state + 18, &m_nullForceTransform, 96);
    parameters[0] = m_length;
    parameters[1] = m_width;
    parameters[2] = m_springCoeff;
    parameters[3] = m_damperCoeff;
}
