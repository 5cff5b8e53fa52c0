#pragma once
// Typed device arrays over DeviceMemory / DeviceMemory2D, API-compatible with the reference's
// kfusion/cuda/device_array.hpp:19-303.
#include <kfusion/cuda/device_memory.hpp>
#include <vector>

namespace kfusion
{
    namespace cuda
    {
        template<class T> class KF_EXPORTS DeviceArray : public DeviceMemory
        {
        public:
            typedef T type;
            enum { elem_size = sizeof(T) };
            DeviceArray() {}
            DeviceArray(size_t size) : DeviceMemory(size * elem_size) {}
            DeviceArray(T *ptr, size_t size) : DeviceMemory(ptr, size * elem_size) {}
            DeviceArray(const DeviceArray& other) : DeviceMemory(other) {}
            DeviceArray& operator = (const DeviceArray& other) { DeviceMemory::operator=(other); return *this; }
            void create(size_t size) { DeviceMemory::create(size * elem_size); }
            void release() { DeviceMemory::release(); }
            void copyTo(DeviceArray& other) const { DeviceMemory::copyTo(other); }
            void upload(const T *host_ptr, size_t size) { DeviceMemory::upload(host_ptr, size * elem_size); }
            void download(T *host_ptr) const { DeviceMemory::download(host_ptr); }
            template<class A> void upload(const std::vector<T, A>& data) { upload(&data[0], data.size()); }
            template<typename A> void download(std::vector<T, A>& data) const { data.resize(size()); if (!data.empty()) download(&data[0]); }
            void swap(DeviceArray& other_arg) { DeviceMemory::swap(other_arg); }
            T* ptr() { return DeviceMemory::ptr<T>(); }
            const T* ptr() const { return DeviceMemory::ptr<T>(); }
            operator T*() { return ptr(); }
            operator const T*() const { return ptr(); }
            size_t size() const { return sizeBytes() / elem_size; }
        };

        template<class T> class KF_EXPORTS DeviceArray2D : public DeviceMemory2D
        {
        public:
            typedef T type;
            enum { elem_size = sizeof(T) };
            DeviceArray2D() {}
            DeviceArray2D(int rows, int cols) : DeviceMemory2D(rows, cols * elem_size) {}
            DeviceArray2D(int rows, int cols, void *data, size_t stepBytes) : DeviceMemory2D(rows, cols * elem_size, data, stepBytes) {}
            DeviceArray2D(const DeviceArray2D& other) : DeviceMemory2D(other) {}
            DeviceArray2D& operator = (const DeviceArray2D& other) { DeviceMemory2D::operator=(other); return *this; }
            void create(int rows, int cols) { DeviceMemory2D::create(rows, cols * elem_size); }
            void release() { DeviceMemory2D::release(); }
            void copyTo(DeviceArray2D& other) const { DeviceMemory2D::copyTo(other); }
            void upload(const void *host_ptr, size_t host_step, int rows, int cols) { DeviceMemory2D::upload(host_ptr, host_step, rows, cols * elem_size); }
            void download(void *host_ptr, size_t host_step) const { DeviceMemory2D::download(host_ptr, host_step); }
            void swap(DeviceArray2D& other_arg) { DeviceMemory2D::swap(other_arg); }
            template<class A> void upload(const std::vector<T, A>& data, int cols) { upload(&data[0], cols * elem_size, (int)(data.size() / cols), cols); }
            template<class A> void download(std::vector<T, A>& data, int& cols) const
            { cols = this->cols(); data.resize((size_t)cols * rows()); if (!data.empty()) download(&data[0], cols * elem_size); }
            T* ptr(int y = 0) { return DeviceMemory2D::ptr<T>(y); }
            const T* ptr(int y = 0) const { return DeviceMemory2D::ptr<T>(y); }
            operator T*() { return ptr(); }
            operator const T*() const { return ptr(); }
            int cols() const { return DeviceMemory2D::colsBytes() / elem_size; }
            int rows() const { return DeviceMemory2D::rows(); }
            size_t elem_step() const { return DeviceMemory2D::step() / elem_size; }
        };
    }
    namespace device
    {
        using kfusion::cuda::DeviceArray;
        using kfusion::cuda::DeviceArray2D;
    }
}
