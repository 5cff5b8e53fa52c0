#pragma once
// Ref-counted device blobs, API- and semantics-compatible with the reference's kfusion/cuda/device_memory.hpp:21-214 /
// src/device_memory.cpp:34-252: create() = cudaMalloc / cudaMallocPitch and is a no-op when the size is unchanged; copies
// share the buffer through an atomically updated host refcount; construction from a user pointer is non-owning.
#include <kfusion/exports.hpp>
#include <kfusion/cuda/kernel_containers.hpp>

namespace kfusion
{
    namespace cuda
    {
        /** prints "KinFu2 error: ..." and exits, like the reference (device_memory.cpp:7-11) */
        KF_EXPORTS void error(const char *error_string, const char *file, const int line, const char *func = "");

        class KF_EXPORTS DeviceMemory
        {
        public:
            DeviceMemory();
            ~DeviceMemory();
            DeviceMemory(size_t sizeBytes_arg);
            DeviceMemory(void *ptr_arg, size_t sizeBytes_arg);
            DeviceMemory(const DeviceMemory& other_arg);
            DeviceMemory& operator=(const DeviceMemory& other_arg);
            void create(size_t sizeBytes_arg);
            void release();
            void copyTo(DeviceMemory& other) const;
            void upload(const void *host_ptr_arg, size_t sizeBytes_arg);
            void download(void *host_ptr_arg) const;
            void swap(DeviceMemory& other_arg);
            template<class T> T* ptr() { return (T*)data_; }
            template<class T> const T* ptr() const { return (const T*)data_; }
            template <class U> operator PtrSz<U>() const { PtrSz<U> r; r.data = (U*)ptr<U>(); r.size = sizeBytes_ / sizeof(U); return r; }
            bool empty() const;
            size_t sizeBytes() const;
        private:
            void *data_;
            size_t sizeBytes_;
            int* refcount_;
        };

        class KF_EXPORTS DeviceMemory2D
        {
        public:
            DeviceMemory2D();
            ~DeviceMemory2D();
            DeviceMemory2D(int rows_arg, int colsBytes_arg);
            DeviceMemory2D(int rows_arg, int colsBytes_arg, void *data_arg, size_t step_arg);
            DeviceMemory2D(const DeviceMemory2D& other_arg);
            DeviceMemory2D& operator=(const DeviceMemory2D& other_arg);
            void create(int rows_arg, int colsBytes_arg);
            void release();
            void copyTo(DeviceMemory2D& other) const;
            void upload(const void *host_ptr_arg, size_t host_step_arg, int rows_arg, int colsBytes_arg);
            void download(void *host_ptr_arg, size_t host_step_arg) const;
            void swap(DeviceMemory2D& other_arg);
            template<class T> T* ptr(int y_arg = 0) { return (T*)((char*)data_ + y_arg * step_); }
            template<class T> const T* ptr(int y_arg = 0) const { return (const T*)((const char*)data_ + y_arg * step_); }
            template <class U> operator PtrStep<U>() const { PtrStep<U> r; r.data = (U*)ptr<U>(); r.step = step_; return r; }
            template <class U> operator PtrStepSz<U>() const
            { PtrStepSz<U> r; r.data = (U*)ptr<U>(); r.step = step_; r.cols = colsBytes_ / sizeof(U); r.rows = rows_; return r; }
            bool empty() const;
            int colsBytes() const;
            int rows() const;
            size_t step() const;
        private:
            void *data_;
            size_t step_;
            int colsBytes_;
            int rows_;
            int* refcount_;
        };
    }
    namespace device
    {
        using kfusion::cuda::DeviceMemory;
        using kfusion::cuda::DeviceMemory2D;
    }
}
